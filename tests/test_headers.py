"""CPU: the parameter-set / picture-header / slice-header writers (wrenc_b200/csrc/headers.cpp) parse back
field by field with an independent spec-level parser (tests/vvc_header_parser.py), and the assembled byte stream splits into the
NAL units reference src/main.rs:223-260,294-389 writes.  (Their total size is pinned against the reference's own output files in
tests/test_reference_pin.py.)"""
import numpy as np
import pytest

import wrenc_b200
import vvc_header_parser as vp

GEOMS = [(352, 288), (1920, 1088), (3840, 2176), (32, 32), (1280, 704)]


@pytest.mark.parametrize("W,H", GEOMS)
@pytest.mark.parametrize("qp", [None, 0, 20, 26, 32, 51, 63])
def test_parameter_sets_parse_back(W, H, qp):
    vps = vp.parse_vps(wrenc_b200.header_rbsp("vps", W, H, qp))
    assert vps["vps_video_parameter_set_id"] == 8 and vps["vps_max_layers_minus1"] == 0 and vps["vps_layer_id"] == [9]
    assert vps["ptl"]["general_profile_idc"] == 0 and vps["ptl"]["gci"]["gci_present_flag"] == 0
    # the reference writes a DPB block the specification does not have for a single-layer VPS (vps.rs: each_layer_is_an_ols
    # = false); a conformant parser reads it as vps_extension_flag = 1 + extension data: ue(0)=1 | ue(8) ue(4) ue(1) | 0 | 0
    assert vps["vps_extension_flag"] == 1
    assert vps["vps_extension_data_bits"] == [0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 0]

    sps = vp.parse_sps(wrenc_b200.header_rbsp("sps", W, H, qp))
    assert sps["sps_seq_parameter_set_id"] == 1 and sps["sps_video_parameter_set_id"] == 8
    assert sps["sps_pic_width_max_in_luma_samples"] == W and sps["sps_pic_height_max_in_luma_samples"] == H
    assert sps["sps_chroma_format_idc"] == 1 and sps["sps_log2_ctu_size_minus5"] == 0 and sps["sps_bitdepth_minus8"] == 0
    assert sps["sps_log2_min_luma_coding_block_size_minus2"] == 0 and sps["sps_qtbtt_dual_tree_intra_flag"] == 0
    assert sps["sps_max_mtt_hierarchy_depth_intra_slice_luma"] == 0 and sps["sps_log2_diff_min_qt_min_cb_intra_slice_luma"] == 0
    assert sps["dpb"] == [dict(max_dec_pic_buffering_minus1=8, max_num_reorder_pics=4, max_latency_increase_plus1=1)]
    assert sps["sps_transform_skip_enabled_flag"] == 1 and sps["sps_log2_transform_skip_max_size_minus2"] == 5 and sps["sps_bdpcm_enabled_flag"] == 0
    assert sps["sps_mts_enabled_flag"] == 1 and sps["sps_explicit_mts_intra_enabled_flag"] == 1 and sps["sps_lfnst_enabled_flag"] == 0
    assert sps["sps_joint_cbcr_enabled_flag"] == 0 and sps["sps_same_qp_table_for_chroma_flag"] == 1
    t = sps["qp_tables"][0]  # identity chroma QP mapping: start 0, 63 points of (in 1, diff 1)
    assert t["start_minus26"] == -26 and t["num_points_minus1"] == 62 and t["points"] == [(0, 1)] * 63
    assert sps["sps_sao_enabled_flag"] == 0 and sps["sps_alf_enabled_flag"] == 0 and sps["sps_lmcs_enabled_flag"] == 0
    assert sps["sps_num_ref_pic_lists"] == [1, 1]
    assert [e["abs_delta_poc_st"] for e in sps["rpls"][0][0]["entries"]] == [0, 2, 3]
    assert [e["strp_entry_sign_flag"] for e in sps["rpls"][0][0]["entries"]] == [1, 1, 1]
    assert [e["strp_entry_sign_flag"] for e in sps["rpls"][1][0]["entries"]] == [0, 0, 0]
    assert sps["sps_cclm_enabled_flag"] == 1 and sps["sps_chroma_horizontal_collocated_flag"] == 0 and sps["sps_chroma_vertical_collocated_flag"] == 0
    assert sps["sps_isp_enabled_flag"] == 0 and sps["sps_mrl_enabled_flag"] == 0 and sps["sps_mip_enabled_flag"] == 0
    assert sps["sps_dep_quant_enabled_flag"] == 1 and sps["sps_sign_data_hiding_enabled_flag"] == 0 and sps["sps_min_qp_prime_ts"] == 0
    assert sps["sps_entropy_coding_sync_enabled_flag"] == 0 and sps["sps_log2_max_pic_order_cnt_lsb_minus4"] == 0

    pps = vp.parse_pps(wrenc_b200.header_rbsp("pps", W, H, qp))
    assert pps["pps_pic_parameter_set_id"] == 1 and pps["pps_seq_parameter_set_id"] == 1
    assert pps["pps_pic_width_in_luma_samples"] == W and pps["pps_pic_height_in_luma_samples"] == H
    assert pps["pps_no_pic_partition_flag"] == 1 and pps["pps_cu_qp_delta_enabled_flag"] == 1
    init_qp = 26 if qp is None else max(qp, 26)
    assert pps["pps_init_qp_minus26"] == init_qp - 26
    assert pps["pps_deblocking_filter_control_present_flag"] == 1 and pps["pps_deblocking_filter_disabled_flag"] == 1
    assert pps["pps_num_ref_idx_default_active_minus1"] == [2, 2]

    for poc in (0, 1, 15, 16, 239):
        ph = vp.parse_ph(wrenc_b200.header_rbsp("ph", W, H, qp, poc), sps, pps)
        assert ph["ph_gdr_or_irap_pic_flag"] == 1 and ph["ph_gdr_pic_flag"] == 0 and ph["ph_inter_slice_allowed_flag"] == 0
        assert ph["ph_pic_parameter_set_id"] == 1 and ph["ph_pic_order_cnt_lsb"] == poc % 16 and ph["ph_cu_qp_delta_subdiv_intra_slice"] == 0
        shb = wrenc_b200.header_rbsp("sh", W, H, qp, poc)
        sh, used = vp.parse_slice_header(shb, sps, pps, ph, 7)
        assert used == len(shb)
        assert sh["sh_no_output_of_prior_pics_flag"] == 0 and sh["sh_dep_quant_used_flag"] == 1
        # SliceQpY = 26 + pps_init_qp_minus26 + sh_qp_delta must be --qp (26 without the flag): H10
        assert 26 + pps["pps_init_qp_minus26"] + sh["sh_qp_delta"] == (26 if qp is None else qp)


def test_byte_stream_layout_and_emulation_prevention():
    rng = np.random.default_rng(7)
    W, H, qp = 96, 64, 23
    sds = []
    for i in range(18):
        sd = bytearray(rng.integers(0, 256, 200 + i, dtype=np.uint8).tobytes())
        sd[10:14] = b"\0\0\0\1"   # must be escaped
        sd[50:53] = b"\0\0\3"
        sd[-1] |= 0x80              # payloads end with a non-zero byte (rbsp stop bit)
        sds.append(bytes(sd))
    stream = wrenc_b200.assemble_vvc(W, H, qp, sds)
    assert stream.startswith(b"\0\0\0\0\0\1")
    nals = vp.split_byte_stream(stream)
    assert [(l, t) for l, t, _, _ in nals[:3]] == [(1, 14), (9, 15), (9, 16)]
    assert all(tid == 1 for _, _, tid, _ in nals)
    assert len(nals) == 3 + 2 * len(sds)
    sps, pps = vp.parse_sps(nals[1][3]), vp.parse_pps(nals[2][3])
    assert vp.parse_vps(nals[0][3])["vps_video_parameter_set_id"] == 8
    for i, sd in enumerate(sds):
        (l0, t0, _, ph_body), (l1, t1, _, body) = nals[3 + 2 * i], nals[4 + 2 * i]
        assert (l0, t0, l1, t1) == (9, 19, 9, 7)
        ph = vp.parse_ph(ph_body, sps, pps)
        assert ph["ph_pic_order_cnt_lsb"] == i % 16
        sh, used = vp.parse_slice_header(body, sps, pps, ph, 7)
        assert sh["sh_qp_delta"] == qp - 26
        assert body[used:] == sd, "slice_data does not survive NAL wrapping + emulation-prevention removal"
    # every NAL unit is exactly what write_nal produces for the raw header bits
    assert wrenc_b200.write_parameter_sets(W, H, qp) == (wrenc_b200.write_nal(wrenc_b200.header_rbsp("vps", W, H, qp), 14, 1)
                                                         + wrenc_b200.write_nal(wrenc_b200.header_rbsp("sps", W, H, qp), 15, 9)
                                                         + wrenc_b200.write_nal(wrenc_b200.header_rbsp("pps", W, H, qp), 16, 9))


def test_header_writer_argument_errors():
    L = wrenc_b200.load_library()
    assert L.wrenc_b200_write_parameter_sets(100, 64, 32, None, 0) == -1      # width not a multiple of 32
    assert L.wrenc_b200_write_parameter_sets(96, 64, 64, None, 0) == -1       # qp out of range
    assert L.wrenc_b200_write_parameter_sets(96, 64, 32, None, 0) < -16       # size query
    assert L.wrenc_b200_write_picture(32, 0, None, 5, None, 0) == -1
