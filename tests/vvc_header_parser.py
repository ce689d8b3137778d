"""Spec-level parser (H.266 | ISO/IEC 23090-3 v1, clauses 7.3.2.3-7.3.2.8, 7.3.3, 7.3.7, 7.3.9, 7.3.10) of the VPS, SPS, PPS, PH
and slice header — test infrastructure, written from the specification's syntax tables and NOT from the reference's writers,
so that it is an independent reading of the bits wrenc_b200/csrc/headers.cpp emits.  Syntax branches the emitted
configuration can never take raise `Untaken` instead of being parsed."""


class Untaken(Exception):
    pass


class Bits:
    def __init__(self, data):
        self.d = bytes(data)
        self.pos = 0

    def u(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | ((self.d[self.pos >> 3] >> (7 - (self.pos & 7))) & 1)
            self.pos += 1
        return v

    def ue(self):
        z = 0
        while self.u(1) == 0:
            z += 1
            assert z < 33
        return (1 << z) - 1 + (self.u(z) if z else 0)

    def se(self):
        k = self.ue()
        return (k + 1) // 2 if k & 1 else -(k // 2)

    def byte_aligned(self):
        return self.pos % 8 == 0

    def align_zero(self):
        while not self.byte_aligned():
            assert self.u(1) == 0, "alignment bit not zero"

    def more_rbsp_data(self):
        """True unless only rbsp_trailing_bits remain (last 1 bit of the payload followed by zeros)."""
        total = len(self.d) * 8
        last = total - 1
        while last >= 0 and ((self.d[last >> 3] >> (7 - (last & 7))) & 1) == 0:
            last -= 1
        return self.pos < last

    def rbsp_trailing_bits(self):
        assert self.u(1) == 1, "rbsp_stop_one_bit"
        self.align_zero()
        assert self.pos == len(self.d) * 8, "bytes after rbsp_trailing_bits"


def general_constraints_info(b):
    g = {"gci_present_flag": b.u(1)}
    if g["gci_present_flag"]:
        raise Untaken("gci fields")
    b.align_zero()
    return g


def profile_tier_level(b, profile_tier_present, max_sublayers_minus1):
    p = {}
    if profile_tier_present:
        p["general_profile_idc"] = b.u(7)
        p["general_tier_flag"] = b.u(1)
    p["general_level_idc"] = b.u(8)
    p["ptl_frame_only_constraint_flag"] = b.u(1)
    p["ptl_multilayer_enabled_flag"] = b.u(1)
    if profile_tier_present:
        p["gci"] = general_constraints_info(b)
    present = [b.u(1) for _ in range(max_sublayers_minus1)]
    b.align_zero()
    p["sublayer_level_idc"] = [b.u(8) for f in present if f]
    if profile_tier_present:
        p["ptl_num_sub_profiles"] = b.u(8)
        p["general_sub_profile_idc"] = [b.u(32) for _ in range(p["ptl_num_sub_profiles"])]
    return p


def dpb_parameters(b, max_sublayers_minus1, sublayer_info_flag):
    out = []
    for _ in range(0 if sublayer_info_flag else max_sublayers_minus1, max_sublayers_minus1 + 1):
        out.append(dict(max_dec_pic_buffering_minus1=b.ue(), max_num_reorder_pics=b.ue(), max_latency_increase_plus1=b.ue()))
    return out


def parse_vps(data):
    """7.3.2.3.  With vps_max_layers_minus1 == 0 the specification infers vps_each_layer_is_an_ols_flag = 1, so the DPB / HRD
    block is absent and whatever follows the profile_tier_level is vps_extension_flag + extension data."""
    b = Bits(data)
    v = {"vps_video_parameter_set_id": b.u(4), "vps_max_layers_minus1": b.u(6), "vps_max_sublayers_minus1": b.u(3)}
    if v["vps_max_layers_minus1"] > 0:
        raise Untaken("multi-layer VPS")
    v["vps_layer_id"] = [b.u(6)]
    # vps_num_ptls_minus1 inferred 0, vps_default_ptl_dpb_hrd_max_tid_flag inferred 1, vps_pt_present_flag[0] inferred 1
    b.align_zero()
    v["ptl"] = profile_tier_level(b, 1, v["vps_max_sublayers_minus1"])
    v["vps_extension_flag"] = b.u(1)
    ext = []
    if v["vps_extension_flag"]:
        while b.more_rbsp_data():
            ext.append(b.u(1))
    v["vps_extension_data_bits"] = ext
    b.rbsp_trailing_bits()
    return v


def ref_pic_list_struct(b, sps, list_idx, rpls_idx):
    r = {"num_ref_entries": b.ue()}
    ltrp_in_header = 0
    if sps["sps_long_term_ref_pics_flag"] and rpls_idx < sps["sps_num_ref_pic_lists"][list_idx] and r["num_ref_entries"] > 0:
        ltrp_in_header = b.u(1)
    r["entries"] = []
    for i in range(r["num_ref_entries"]):
        e = {}
        ilrp = b.u(1) if sps["sps_inter_layer_prediction_enabled_flag"] else 0
        if not ilrp:
            st = b.u(1) if sps["sps_long_term_ref_pics_flag"] else 1
            if st:
                e["abs_delta_poc_st"] = b.ue()
                a = e["abs_delta_poc_st"] if ((sps["sps_weighted_pred_flag"] or sps["sps_weighted_bipred_flag"]) and i != 0) else e["abs_delta_poc_st"] + 1
                if a > 0:
                    e["strp_entry_sign_flag"] = b.u(1)
            elif not ltrp_in_header:
                e["rpls_poc_lsb_lt"] = b.u(sps["sps_log2_max_pic_order_cnt_lsb_minus4"] + 4)
        else:
            e["ilrp_idx"] = b.ue()
        r["entries"].append(e)
    return r


def parse_sps(data):
    """7.3.2.4"""
    b = Bits(data)
    s = {}

    def f(name, n=1):
        s[name] = b.u(n)
        return s[name]

    def ue(name):
        s[name] = b.ue()
        return s[name]

    f("sps_seq_parameter_set_id", 4)
    f("sps_video_parameter_set_id", 4)
    f("sps_max_sublayers_minus1", 3)
    f("sps_chroma_format_idc", 2)
    f("sps_log2_ctu_size_minus5", 2)
    if f("sps_ptl_dpb_hrd_params_present_flag"):
        s["ptl"] = profile_tier_level(b, 1, s["sps_max_sublayers_minus1"])
    f("sps_gdr_enabled_flag")
    if f("sps_ref_pic_resampling_enabled_flag"):
        f("sps_res_change_in_clvs_allowed_flag")
    ue("sps_pic_width_max_in_luma_samples")
    ue("sps_pic_height_max_in_luma_samples")
    if f("sps_conformance_window_flag"):
        raise Untaken("conformance window")
    if f("sps_subpic_info_present_flag"):
        raise Untaken("subpictures")
    ue("sps_bitdepth_minus8")
    f("sps_entropy_coding_sync_enabled_flag")
    f("sps_entry_point_offsets_present_flag")
    f("sps_log2_max_pic_order_cnt_lsb_minus4", 4)
    if f("sps_poc_msb_cycle_flag"):
        ue("sps_poc_msb_cycle_len_minus1")
    f("sps_num_extra_ph_bytes", 2)
    s["sps_extra_ph_bit_present_flag"] = [b.u(1) for _ in range(8 * s["sps_num_extra_ph_bytes"])]
    f("sps_num_extra_sh_bytes", 2)
    s["sps_extra_sh_bit_present_flag"] = [b.u(1) for _ in range(8 * s["sps_num_extra_sh_bytes"])]
    if s["sps_ptl_dpb_hrd_params_present_flag"]:
        sub = f("sps_sublayer_dpb_params_flag") if s["sps_max_sublayers_minus1"] > 0 else 0
        s["dpb"] = dpb_parameters(b, s["sps_max_sublayers_minus1"], sub)
    ue("sps_log2_min_luma_coding_block_size_minus2")
    f("sps_partition_constraints_override_enabled_flag")
    ue("sps_log2_diff_min_qt_min_cb_intra_slice_luma")
    if ue("sps_max_mtt_hierarchy_depth_intra_slice_luma") != 0:
        ue("sps_log2_diff_max_bt_min_qt_intra_slice_luma")
        ue("sps_log2_diff_max_tt_min_qt_intra_slice_luma")
    s["sps_qtbtt_dual_tree_intra_flag"] = b.u(1) if s["sps_chroma_format_idc"] != 0 else 0
    if s["sps_qtbtt_dual_tree_intra_flag"]:
        ue("sps_log2_diff_min_qt_min_cb_intra_slice_chroma")
        if ue("sps_max_mtt_hierarchy_depth_intra_slice_chroma") != 0:
            ue("sps_log2_diff_max_bt_min_qt_intra_slice_chroma")
            ue("sps_log2_diff_max_tt_min_qt_intra_slice_chroma")
    ue("sps_log2_diff_min_qt_min_cb_inter_slice")
    if ue("sps_max_mtt_hierarchy_depth_inter_slice") != 0:
        ue("sps_log2_diff_max_bt_min_qt_inter_slice")
        ue("sps_log2_diff_max_tt_min_qt_inter_slice")
    if (1 << (s["sps_log2_ctu_size_minus5"] + 5)) > 32:
        f("sps_max_luma_transform_size_64_flag")
    if f("sps_transform_skip_enabled_flag"):
        ue("sps_log2_transform_skip_max_size_minus2")
        f("sps_bdpcm_enabled_flag")
    if f("sps_mts_enabled_flag"):
        f("sps_explicit_mts_intra_enabled_flag")
        f("sps_explicit_mts_inter_enabled_flag")
    f("sps_lfnst_enabled_flag")
    if s["sps_chroma_format_idc"] != 0:
        f("sps_joint_cbcr_enabled_flag")
        f("sps_same_qp_table_for_chroma_flag")
        n_tab = 1 if s["sps_same_qp_table_for_chroma_flag"] else (3 if s["sps_joint_cbcr_enabled_flag"] else 2)
        s["qp_tables"] = []
        for _ in range(n_tab):
            t = dict(start_minus26=b.se(), num_points_minus1=b.ue(), points=[])
            for _ in range(t["num_points_minus1"] + 1):
                t["points"].append((b.ue(), b.ue()))  # (delta_qp_in_val_minus1, delta_qp_diff_val)
            s["qp_tables"].append(t)
    f("sps_sao_enabled_flag")
    if f("sps_alf_enabled_flag") and s["sps_chroma_format_idc"] != 0:
        f("sps_ccalf_enabled_flag")
    f("sps_lmcs_enabled_flag")
    f("sps_weighted_pred_flag")
    f("sps_weighted_bipred_flag")
    f("sps_long_term_ref_pics_flag")
    s["sps_inter_layer_prediction_enabled_flag"] = b.u(1) if s["sps_video_parameter_set_id"] > 0 else 0
    f("sps_idr_rpl_present_flag")
    f("sps_rpl1_same_as_rpl0_flag")
    s["sps_num_ref_pic_lists"] = [0, 0]
    s["rpls"] = [[], []]
    for i in range(1 if s["sps_rpl1_same_as_rpl0_flag"] else 2):
        s["sps_num_ref_pic_lists"][i] = b.ue()
        for j in range(s["sps_num_ref_pic_lists"][i]):
            s["rpls"][i].append(ref_pic_list_struct(b, s, i, j))
    f("sps_ref_wraparound_enabled_flag")
    if f("sps_temporal_mvp_enabled_flag"):
        f("sps_sbtmvp_enabled_flag")
    f("sps_amvr_enabled_flag")
    if f("sps_bdof_enabled_flag"):
        f("sps_bdof_control_present_in_ph_flag")
    f("sps_smvd_enabled_flag")
    if f("sps_dmvr_enabled_flag"):
        f("sps_dmvr_control_present_in_ph_flag")
    if f("sps_mmvd_enabled_flag"):
        f("sps_mmvd_fullpel_only_enabled_flag")
    max_merge = 6 - ue("sps_six_minus_max_num_merge_cand")
    f("sps_sbt_enabled_flag")
    if f("sps_affine_enabled_flag"):
        raise Untaken("affine")
    f("sps_bcw_enabled_flag")
    f("sps_ciip_enabled_flag")
    if max_merge >= 2:
        if f("sps_gpm_enabled_flag") and max_merge >= 3:
            ue("sps_max_num_merge_cand_minus_max_num_gpm_cand")
    ue("sps_log2_parallel_merge_level_minus2")
    f("sps_isp_enabled_flag")
    f("sps_mrl_enabled_flag")
    f("sps_mip_enabled_flag")
    if s["sps_chroma_format_idc"] != 0:
        f("sps_cclm_enabled_flag")
    if s["sps_chroma_format_idc"] == 1:
        f("sps_chroma_horizontal_collocated_flag")
        f("sps_chroma_vertical_collocated_flag")
    f("sps_palette_enabled_flag")
    if s["sps_chroma_format_idc"] == 3 and not s.get("sps_max_luma_transform_size_64_flag", 0):
        f("sps_act_enabled_flag")
    if s["sps_transform_skip_enabled_flag"] or s["sps_palette_enabled_flag"]:
        ue("sps_min_qp_prime_ts")
    if f("sps_ibc_enabled_flag"):
        ue("sps_six_minus_max_num_ibc_merge_cand")
    if f("sps_ladf_enabled_flag"):
        raise Untaken("ladf")
    if f("sps_explicit_scaling_list_enabled_flag"):
        raise Untaken("scaling lists")
    f("sps_dep_quant_enabled_flag")
    f("sps_sign_data_hiding_enabled_flag")
    if f("sps_virtual_boundaries_enabled_flag"):
        raise Untaken("virtual boundaries")
    if s["sps_ptl_dpb_hrd_params_present_flag"]:
        if f("sps_timing_hrd_params_present_flag"):
            raise Untaken("hrd")
    f("sps_field_seq_flag")
    if f("sps_vui_parameters_present_flag"):
        raise Untaken("vui")
    if f("sps_extension_flag"):
        raise Untaken("sps extension")
    b.rbsp_trailing_bits()
    return s


def parse_pps(data):
    """7.3.2.5"""
    b = Bits(data)
    p = {}

    def f(name, n=1):
        p[name] = b.u(n)
        return p[name]

    f("pps_pic_parameter_set_id", 6)
    f("pps_seq_parameter_set_id", 4)
    f("pps_mixed_nalu_types_in_pic_flag")
    p["pps_pic_width_in_luma_samples"] = b.ue()
    p["pps_pic_height_in_luma_samples"] = b.ue()
    if f("pps_conformance_window_flag"):
        raise Untaken("conformance window")
    if f("pps_scaling_window_explicit_signalling_flag"):
        raise Untaken("scaling window")
    f("pps_output_flag_present_flag")
    f("pps_no_pic_partition_flag")
    if f("pps_subpic_id_mapping_present_flag"):
        raise Untaken("subpic id mapping")
    if not p["pps_no_pic_partition_flag"]:
        raise Untaken("picture partitioning")
    f("pps_cabac_init_present_flag")
    p["pps_num_ref_idx_default_active_minus1"] = [b.ue(), b.ue()]
    f("pps_rpl1_idx_present_flag")
    f("pps_weighted_pred_flag")
    f("pps_weighted_bipred_flag")
    if f("pps_ref_wraparound_enabled_flag"):
        p["pps_pic_width_minus_wraparound_offset"] = b.ue()
    p["pps_init_qp_minus26"] = b.se()
    f("pps_cu_qp_delta_enabled_flag")
    if f("pps_chroma_tool_offsets_present_flag"):
        raise Untaken("chroma tool offsets")
    p["pps_deblocking_filter_override_enabled_flag"] = 0
    p["pps_deblocking_filter_disabled_flag"] = 0
    if f("pps_deblocking_filter_control_present_flag"):
        f("pps_deblocking_filter_override_enabled_flag")
        f("pps_deblocking_filter_disabled_flag")
        if not p["pps_no_pic_partition_flag"] and p["pps_deblocking_filter_override_enabled_flag"]:
            f("pps_dbf_info_in_ph_flag")
        if not p["pps_deblocking_filter_disabled_flag"]:
            p["pps_luma_beta_offset_div2"] = b.se()
            p["pps_luma_tc_offset_div2"] = b.se()
    # pps_no_pic_partition_flag = 1: pps_rpl/sao/alf/wp/qp_delta_info_in_ph_flag and pps_dbf_info_in_ph_flag inferred 0, pps_rect_slice_flag inferred 1
    for k in ("pps_rpl_info_in_ph_flag", "pps_sao_info_in_ph_flag", "pps_alf_info_in_ph_flag", "pps_wp_info_in_ph_flag", "pps_qp_delta_info_in_ph_flag"):
        p[k] = 0
    p.setdefault("pps_dbf_info_in_ph_flag", 0)
    f("pps_picture_header_extension_present_flag")
    f("pps_slice_header_extension_present_flag")
    if f("pps_extension_flag"):
        raise Untaken("pps extension")
    b.rbsp_trailing_bits()
    return p


def picture_header_structure(b, sps, pps):
    """7.3.2.8"""
    h = {}

    def f(name, n=1):
        h[name] = b.u(n)
        return h[name]

    f("ph_gdr_or_irap_pic_flag")
    f("ph_non_ref_pic_flag")
    h["ph_gdr_pic_flag"] = b.u(1) if h["ph_gdr_or_irap_pic_flag"] else 0
    h["ph_intra_slice_allowed_flag"] = 1
    if f("ph_inter_slice_allowed_flag"):
        f("ph_intra_slice_allowed_flag")
    h["ph_pic_parameter_set_id"] = b.ue()
    f("ph_pic_order_cnt_lsb", sps["sps_log2_max_pic_order_cnt_lsb_minus4"] + 4)
    if h["ph_gdr_pic_flag"]:
        h["ph_recovery_poc_cnt"] = b.ue()
    h["ph_extra_bit"] = [b.u(1) for x in sps["sps_extra_ph_bit_present_flag"] if x]
    if sps["sps_poc_msb_cycle_flag"]:
        raise Untaken("poc msb")
    if sps["sps_alf_enabled_flag"] and pps["pps_alf_info_in_ph_flag"]:
        raise Untaken("alf")
    if sps["sps_lmcs_enabled_flag"]:
        raise Untaken("lmcs")
    # sps_explicit_scaling_list / virtual boundaries: parse_sps raises when enabled
    if pps["pps_output_flag_present_flag"] and not h["ph_non_ref_pic_flag"]:
        f("ph_pic_output_flag")
    if pps["pps_rpl_info_in_ph_flag"]:
        raise Untaken("rpl in ph")
    h["ph_partition_constraints_override_flag"] = b.u(1) if sps["sps_partition_constraints_override_enabled_flag"] else 0
    if h["ph_intra_slice_allowed_flag"]:
        if h["ph_partition_constraints_override_flag"]:
            raise Untaken("partition override")
        if pps["pps_cu_qp_delta_enabled_flag"]:
            h["ph_cu_qp_delta_subdiv_intra_slice"] = b.ue()
        # pps_cu_chroma_qp_offset_list_enabled_flag = 0 (no chroma tool offsets)
    if h["ph_inter_slice_allowed_flag"]:
        raise Untaken("inter slices")
    if pps["pps_qp_delta_info_in_ph_flag"]:
        h["ph_qp_delta"] = b.se()
    if sps["sps_joint_cbcr_enabled_flag"]:
        f("ph_joint_cbcr_sign_flag")
    if sps["sps_sao_enabled_flag"] and pps["pps_sao_info_in_ph_flag"]:
        raise Untaken("sao")
    if pps["pps_dbf_info_in_ph_flag"]:
        raise Untaken("dbf in ph")
    if pps["pps_picture_header_extension_present_flag"]:
        raise Untaken("ph extension")
    return h


def parse_ph(data, sps, pps):
    b = Bits(data)
    h = picture_header_structure(b, sps, pps)
    b.rbsp_trailing_bits()
    return h


def parse_slice_header(data, sps, pps, ph, nal_unit_type):
    """7.3.7 up to and including byte_alignment(); returns (fields, header length in bytes)."""
    b = Bits(data)
    h = {}
    if b.u(1):
        raise Untaken("ph in sh")
    # sps_subpic_info_present_flag = 0; pps_rect_slice_flag inferred 1 with NumSlicesInSubpic = 1: no sh_slice_address;
    # NumExtraShBits = 0; no sh_num_tiles_in_slice_minus1
    assert not any(sps["sps_extra_sh_bit_present_flag"])
    h["sh_slice_type"] = b.ue() if ph["ph_inter_slice_allowed_flag"] else 2
    if nal_unit_type in (7, 8, 9, 10):  # IDR_W_RADL, IDR_N_LP, CRA_NUT, GDR_NUT
        h["sh_no_output_of_prior_pics_flag"] = b.u(1)
    if sps["sps_alf_enabled_flag"] and not pps["pps_alf_info_in_ph_flag"]:
        raise Untaken("alf")
    if not pps["pps_rpl_info_in_ph_flag"] and ((nal_unit_type not in (7, 8)) or sps["sps_idr_rpl_present_flag"]):
        raise Untaken("ref pic lists")
    assert h["sh_slice_type"] == 2
    if not pps["pps_qp_delta_info_in_ph_flag"]:
        h["sh_qp_delta"] = b.se()
    # pps_slice_chroma_qp_offsets_present_flag = 0, pps_cu_chroma_qp_offset_list_enabled_flag = 0
    if sps["sps_sao_enabled_flag"] and not pps["pps_sao_info_in_ph_flag"]:
        raise Untaken("sao")
    h["sh_deblocking_params_present_flag"] = 0
    if pps["pps_deblocking_filter_override_enabled_flag"] and not pps["pps_dbf_info_in_ph_flag"]:
        h["sh_deblocking_params_present_flag"] = b.u(1)
    if h["sh_deblocking_params_present_flag"]:
        raise Untaken("deblocking params")
    h["sh_dep_quant_used_flag"] = b.u(1) if sps["sps_dep_quant_enabled_flag"] else 0
    h["sh_sign_data_hiding_used_flag"] = 0
    if sps["sps_sign_data_hiding_enabled_flag"] and not h["sh_dep_quant_used_flag"]:
        h["sh_sign_data_hiding_used_flag"] = b.u(1)
    if sps["sps_transform_skip_enabled_flag"] and not h["sh_dep_quant_used_flag"] and not h["sh_sign_data_hiding_used_flag"]:
        h["sh_ts_residual_coding_disabled_flag"] = b.u(1)
    if pps["pps_slice_header_extension_present_flag"]:
        raise Untaken("sh extension")
    # NumEntryPoints = 0 (sps_entry_point_offsets_present_flag = 0)
    assert b.u(1) == 1, "byte_alignment(): alignment_bit_equal_to_one"
    b.align_zero()
    return h, b.pos // 8


def split_byte_stream(stream):
    """Annex B: split at start codes; returns [(nuh_layer_id, nal_unit_type, temporal_id_plus1, rbsp bytes with emulation
    prevention removed)]."""
    out = []
    i, n = 0, len(stream)
    starts = []
    while i + 2 < n:
        if stream[i] == 0 and stream[i + 1] == 0 and stream[i + 2] == 1:
            starts.append(i + 3)
            i += 3
        else:
            i += 1
    for k, s in enumerate(starts):
        e = n if k + 1 == len(starts) else starts[k + 1] - 3
        nal = stream[s:e]
        while k + 1 < len(starts) and nal and nal[-1] == 0:  # trailing_zero_8bits / leading zeros of the next start code
            nal = nal[:-1]
        hdr0, hdr1 = nal[0], nal[1]
        assert hdr0 >> 6 == 0, "forbidden_zero_bit / nuh_reserved_zero_bit"
        body = bytearray()
        z = 0
        for x in nal[2:]:
            if z >= 2 and x == 3:
                z = 0
                continue
            body.append(x)
            z = z + 1 if x == 0 else 0
        out.append((hdr0 & 63, hdr1 >> 3, hdr1 & 7, bytes(body)))
    return out
