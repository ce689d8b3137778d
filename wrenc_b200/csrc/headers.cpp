// headers.cpp — the parameter-set / picture-header / slice-header bits of the one configuration hjmkt/wrenc emits
// (SURVEY.md §8 row f-1).  Pure host code, no GPU: with wrenc_b200_write_nal it turns the slice_data() bytes the device
// coder returns into a complete .vvc byte stream, NAL unit by NAL unit, as reference src/main.rs:223-260,294-389 does.
//
// The reference serialises plain structs with hard-coded defaults through ~450 conditional writes; for its fixed
// configuration only the branches below are taken.  Every element is written with the value and the order of the
// reference's encoder (file:line cited per group); the variable inputs are width, height, --qp and the picture index.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/wrenc_b200.h"

namespace {

// MSB-first bit string: the observable behaviour of reference src/bins.rs (push_bin / push_bins_with_size / byte_align)
// and of BoolCoder::encode_{un,}signed_exp_golomb (src/bool_coder.rs:62-84).
struct BitString {
    std::vector<uint8_t> bytes;
    int nbits = 0;  // bits used in the last byte (0 = byte aligned)
    void u(uint64_t v, int n) {
        for (int i = n - 1; i >= 0; i--) {
            if (nbits == 0) bytes.push_back(0);
            bytes.back() |= (uint8_t)(((v >> i) & 1) << (7 - nbits));
            nbits = (nbits + 1) & 7;
        }
    }
    void flag(bool b) { u(b ? 1 : 0, 1); }
    void ue(uint64_t v) {  // bool_coder.rs:62-74
        int n = 0;
        while (((v + 1) >> (n + 1)) != 0) n++;
        u(0, n);
        u(1, 1);
        u(v + 1 - ((uint64_t)1 << n), n);
    }
    void se(int64_t v) {  // bool_coder.rs:76-84
        if (v == 0) { ue(0); return; }
        uint64_t a = (uint64_t)(v < 0 ? -v : v);
        ue((a - 1) * 2 + 1 + (v < 0 ? 1 : 0));
    }
    void byte_align() { nbits = 0; }  // bins.rs:126-132 (zero padding)
    void trailing() { flag(true); byte_align(); }  // rbsp_stop_one_bit + alignment
};

// dpb_parameters( ) with DpbParameter::new() (src/dpb.rs:11-20: 8, 4, 1), one sub-layer (src/dpbp_encoder.rs:26-50)
void dpb_parameters(BitString &b) {
    b.ue(8);  // max_dec_pic_buffering
    b.ue(4);  // max_num_reorder_pics
    b.ue(1);  // max_latency_increase
}

// profile_tier_level( ) of ProfileTierLevel::new(true) (src/ptl.rs:17-32: everything zero, no GCI) for one sub-layer
// (src/ptl_encoder.rs:26-69; general_constraints_info( ) = gci_present_flag 0 + byte alignment, src/gci_encoder.rs:24-109)
void profile_tier_level(BitString &b) {
    b.u(0, 7);       // general_profile_idc
    b.flag(false);   // general_tier_flag
    b.u(0, 8);       // general_level_idc
    b.flag(false);   // ptl_frame_only_constraint_flag
    b.flag(false);   // ptl_multilayer_enabled_flag
    b.flag(false);   // gci_present_flag
    b.byte_align();  // gci alignment (gci_encoder.rs:108)
    b.byte_align();  // ptl alignment (ptl_encoder.rs:54); no sub-layer level flags for max_num_sublayers = 1
    b.u(0, 8);       // ptl_num_sub_profiles
}

// src/vps_encoder.rs:29-279 with VideoParameterSet::new(8, ..) (src/vps.rs:82-117): one layer (id 9), one sub-layer,
// each_layer_is_an_ols = false (so the DPB block is written), total_num_olss = 1, num_multi_layer_olss = 0
// (src/encoder_context.rs:360-546), no HRD, no extension.
std::vector<uint8_t> vps_rbsp() {
    BitString b;
    b.u(8, 4);       // vps_video_parameter_set_id (main.rs:223)
    b.u(0, 6);       // vps_max_layers_minus1
    b.u(0, 3);       // vps_max_sublayers_minus1
    b.u(9, 6);       // vps_layer_id[0] (VpsLayer::new(9))
    b.byte_align();  // vps_encoder.rs:114 (vps_num_ptls = 1: no pt_present / max_tid elements)
    profile_tier_level(b);
    b.ue(0);         // vps_num_dpb_params_minus1 (dpb_parameters.len() = 1), written because !each_layer_is_an_ols
    dpb_parameters(b);
    b.flag(false);   // vps_timing_hrd_params_present_flag
    b.flag(false);   // vps_extension_flag
    b.trailing();
    return b.bytes;
}

// src/sps_encoder.rs:29-665 with SequenceParameterSet::new(1, 8, W, H, 8) (src/sps.rs:228-347),
// PartitionConstraints::new() (src/partition.rs:22-43), QpTable::new(8, 63, 0) (src/sps.rs:36-57),
// RefPicList::new(lx) (src/reference_picture.rs:13-55), ctb_size_y = 32, max_num_merge_cand = 6.
std::vector<uint8_t> sps_rbsp(int W, int H) {
    BitString b;
    b.u(1, 4);       // sps_seq_parameter_set_id
    b.u(8, 4);       // sps_video_parameter_set_id
    b.u(0, 3);       // sps_max_sublayers_minus1
    b.u(1, 2);       // sps_chroma_format_idc (4:2:0)
    b.u(0, 2);       // sps_log2_ctu_size_minus5 (CTU 32)
    b.flag(true);    // sps_ptl_dpb_hrd_params_present_flag
    profile_tier_level(b);
    b.flag(false);   // sps_gdr_enabled_flag
    b.flag(false);   // sps_ref_pic_resampling_enabled_flag
    b.ue((uint64_t)W);  // sps_pic_width_max_in_luma_samples
    b.ue((uint64_t)H);  // sps_pic_height_max_in_luma_samples
    b.flag(false);   // sps_conformance_window_flag
    b.flag(false);   // sps_subpic_info_present_flag
    b.ue(0);         // sps_bitdepth_minus8
    b.flag(false);   // sps_entropy_coding_sync_enabled_flag
    b.flag(false);   // sps_entry_point_offsets_present_flag
    b.u(0, 4);       // sps_log2_max_pic_order_cnt_lsb_minus4
    b.flag(false);   // sps_poc_msb_cycle_flag
    b.u(0, 2);       // sps_num_extra_ph_bytes
    b.u(0, 2);       // sps_num_extra_sh_bytes
    dpb_parameters(b);  // sps_encoder.rs:226-241 (max_sublayers = 1: no sublayer flag)
    b.ue(0);         // sps_log2_min_luma_coding_block_size_minus2 (min CB 4)
    b.flag(false);   // sps_partition_constraints_override_enabled_flag
    b.ue(0);         // sps_log2_diff_min_qt_min_cb_intra_slice_luma
    b.ue(0);         // sps_max_mtt_hierarchy_depth_intra_slice_luma
    b.flag(false);   // sps_qtbtt_dual_tree_intra_flag
    b.ue(0);         // sps_log2_diff_min_qt_min_cb_inter_slice
    b.ue(0);         // sps_max_mtt_hierarchy_depth_inter_slice
                     // ctb_size_y = 32: no sps_max_luma_transform_size_64_flag (sps_encoder.rs:366-369)
    b.flag(true);    // sps_transform_skip_enabled_flag
    b.ue(5);         // sps_log2_transform_skip_max_size_minus2 (the reference writes log2_transform_skip_max_size = 5 as is)
    b.flag(false);   // sps_bdpcm_enabled_flag
    b.flag(true);    // sps_mts_enabled_flag
    b.flag(true);    // sps_explicit_mts_intra_enabled_flag
    b.flag(true);    // sps_explicit_mts_inter_enabled_flag
    b.flag(false);   // sps_lfnst_enabled_flag
    b.flag(false);   // sps_joint_cbcr_enabled_flag
    b.flag(true);    // sps_same_qp_table_for_chroma_flag
    b.se(0 - 26);    // sps_qp_table_start_minus26[0]
    b.ue(63 - 1);    // sps_num_points_in_qp_table_minus1[0]
    for (int j = 0; j < 63; j++) {
        b.ue(1 - 1);  // sps_delta_qp_in_val_minus1
        b.ue(1);      // sps_delta_qp_diff_val
    }
    b.flag(false);   // sps_sao_enabled_flag
    b.flag(false);   // sps_alf_enabled_flag
    b.flag(false);   // sps_lmcs_enabled_flag
    b.flag(false);   // sps_weighted_pred_flag
    b.flag(false);   // sps_weighted_bipred_flag
    b.flag(false);   // sps_long_term_ref_pics_flag
    b.flag(false);   // sps_inter_layer_prediction_enabled_flag (vps id 8 > 0)
    b.flag(false);   // sps_idr_rpl_present_flag
    b.flag(false);   // sps_rpl1_same_as_rpl0_flag
    for (int lx = 0; lx < 2; lx++) {  // sps_encoder.rs:400-418, src/rpl_encoder.rs:75-124
        b.ue(1);                      // sps_num_ref_pic_lists[lx]
        b.ue(3);                      // num_ref_entries
        const int abs_delta_poc_st[3] = {0, 2, 3};
        for (int i = 0; i < 3; i++) {
            b.ue((uint64_t)abs_delta_poc_st[i]);  // abs_delta_poc_st (st_ref_pic_flag inferred 1)
            b.flag(lx == 0);                       // strp_entry_sign_flag (abs_delta_poc_st + 1 > 0 always)
        }
    }
    b.flag(false);   // sps_ref_wraparound_enabled_flag
    b.flag(false);   // sps_temporal_mvp_enabled_flag
    b.flag(false);   // sps_amvr_enabled_flag
    b.flag(false);   // sps_bdof_enabled_flag
    b.flag(false);   // sps_smvd_enabled_flag
    b.flag(false);   // sps_dmvr_enabled_flag
    b.flag(false);   // sps_mmvd_enabled_flag
    b.ue(0);         // sps_six_minus_max_num_merge_cand
    b.flag(false);   // sps_sbt_enabled_flag
    b.flag(false);   // sps_affine_enabled_flag
    b.flag(false);   // sps_bcw_enabled_flag
    b.flag(false);   // sps_ciip_enabled_flag
    b.flag(false);   // sps_gpm_enabled_flag (max_num_merge_cand = 6 >= 2)
    b.ue(0);         // sps_log2_parallel_merge_level_minus2
    b.flag(false);   // sps_isp_enabled_flag
    b.flag(false);   // sps_mrl_enabled_flag
    b.flag(false);   // sps_mip_enabled_flag
    b.flag(true);    // sps_cclm_enabled_flag
    b.flag(false);   // sps_chroma_horizontal_collocated_flag
    b.flag(false);   // sps_chroma_vertical_collocated_flag
    b.flag(false);   // sps_palette_enabled_flag
    b.ue(0);         // sps_min_qp_prime_ts (transform skip enabled)
    b.flag(false);   // sps_ibc_enabled_flag
    b.flag(false);   // sps_ladf_enabled_flag
    b.flag(false);   // sps_explicit_scaling_list_enabled_flag
    b.flag(true);    // sps_dep_quant_enabled_flag
    b.flag(false);   // sps_sign_data_hiding_enabled_flag
    b.flag(false);   // sps_virtual_boundaries_enabled_flag
    b.flag(false);   // sps_timing_hrd_params_present_flag
    b.flag(false);   // sps_field_seq_flag
    b.flag(false);   // sps_vui_parameters_present_flag
    b.flag(false);   // sps_extension_flag
    b.trailing();
    return b.bytes;
}

// init_qp = max(--qp, 26), 26 without --qp (src/pps.rs:183-187)
inline int init_qp_of(int qp) { return qp < 0 ? 26 : (qp > 26 ? qp : 26); }

// src/pps_encoder.rs:29-350 with PictureParameterSet::new(1, &sps, qp) (src/pps.rs:148-197): no picture partitioning,
// deblocking control present with the filter disabled.
std::vector<uint8_t> pps_rbsp(int W, int H, int qp) {
    BitString b;
    b.u(1, 6);       // pps_pic_parameter_set_id
    b.u(1, 4);       // pps_seq_parameter_set_id
    b.flag(false);   // pps_mixed_nalu_types_in_pic_flag
    b.ue((uint64_t)W);  // pps_pic_width_in_luma_samples
    b.ue((uint64_t)H);  // pps_pic_height_in_luma_samples
    b.flag(false);   // pps_conformance_window_flag
    b.flag(false);   // pps_scaling_window_explicit_signalling_flag
    b.flag(false);   // pps_output_flag_present_flag
    b.flag(true);    // pps_no_pic_partition_flag
    b.flag(false);   // pps_subpic_id_mapping_present_flag
    b.flag(false);   // pps_cabac_init_present_flag
    b.ue(3 - 1);     // pps_num_ref_idx_default_active_minus1[0]
    b.ue(3 - 1);     // pps_num_ref_idx_default_active_minus1[1]
    b.flag(false);   // pps_rpl1_idx_present_flag
    b.flag(false);   // pps_weighted_pred_flag
    b.flag(false);   // pps_weighted_bipred_flag
    b.flag(false);   // pps_ref_wraparound_enabled_flag
    b.se(init_qp_of(qp) - 26);  // pps_init_qp_minus26
    b.flag(true);    // pps_cu_qp_delta_enabled_flag
    b.flag(false);   // pps_chroma_tool_offsets_present_flag
    b.flag(true);    // pps_deblocking_filter_control_present_flag
    b.flag(false);   // pps_deblocking_filter_override_enabled_flag
    b.flag(true);    // pps_deblocking_filter_disabled_flag
    b.flag(false);   // pps_picture_header_extension_present_flag
    b.flag(false);   // pps_slice_header_extension_present_flag
    b.flag(false);   // pps_extension_flag
    b.trailing();
    return b.bytes;
}

// src/ph_encoder.rs:29-459 with PictureHeader::new(&pps, true, poc) (src/picture_header.rs:88-141)
std::vector<uint8_t> ph_rbsp(uint64_t picture_index) {
    BitString b;
    b.flag(true);    // ph_gdr_or_irap_pic_flag
    b.flag(false);   // ph_non_ref_pic_flag
    b.flag(false);   // ph_gdr_pic_flag
    b.flag(false);   // ph_inter_slice_allowed_flag
    b.ue(1);         // ph_pic_parameter_set_id
    b.u(picture_index & 15, 4);  // ph_pic_order_cnt_lsb = poc & 0b1111 (picture_header.rs:97)
    b.ue(0);         // ph_cu_qp_delta_subdiv_intra_slice (pps_cu_qp_delta_enabled_flag)
    b.trailing();
    return b.bytes;
}

// src/slice_encoder.rs:32-341 with SliceHeader::new (src/slice_header.rs:63-123): sh_qp_delta = --qp - init_qp
// (ph_encoder.rs:395-397 resets slice_qp_y to init_qp before every SliceHeader::new)
std::vector<uint8_t> slice_header_bits(int qp) {
    BitString b;
    b.flag(false);   // sh_picture_header_in_slice_header_flag
    b.flag(false);   // sh_no_output_of_prior_pics_flag (IDR_W_RADL)
    b.se(qp < 0 ? 0 : qp - init_qp_of(qp));  // sh_qp_delta
    b.flag(true);    // sh_dep_quant_used_flag
    b.flag(true);    // byte_alignment( ): alignment_bit_equal_to_one
    b.byte_align();
    return b.bytes;
}

int64_t emit(std::vector<uint8_t> &dst, int layer, int type, const std::vector<uint8_t> &payload) {
    int64_t need = -wrenc_b200_write_nal(layer, type, 0, payload.data(), payload.size(), nullptr, 0);
    if (need <= 0) return WRENC_B200_EINVAL;
    size_t at = dst.size();
    dst.resize(at + (size_t)need);
    return wrenc_b200_write_nal(layer, type, 0, payload.data(), payload.size(), dst.data() + at, (size_t)need);
}

int64_t hand_out(const std::vector<uint8_t> &v, uint8_t *out, size_t cap) {
    if (!out || cap < v.size()) return -(int64_t)v.size();
    memcpy(out, v.data(), v.size());
    return (int64_t)v.size();
}

bool geometry_ok(int W, int H, int qp) { return W > 0 && H > 0 && W % 32 == 0 && H % 32 == 0 && qp >= -1 && qp <= 63; }

}  // namespace

extern "C" {

// VPS (nuh_layer_id 1), SPS and PPS (nuh_layer_id 9) byte-stream NAL units, main.rs:223-260
int64_t wrenc_b200_write_parameter_sets(int32_t width, int32_t height, int32_t qp, uint8_t *out, size_t cap) {
    if (!geometry_ok(width, height, qp)) return WRENC_B200_EINVAL;
    std::vector<uint8_t> v;
    if (emit(v, 1, 14 /*VPS_NUT*/, vps_rbsp()) < 0) return WRENC_B200_EINVAL;
    if (emit(v, 9, 15 /*SPS_NUT*/, sps_rbsp(width, height)) < 0) return WRENC_B200_EINVAL;
    if (emit(v, 9, 16 /*PPS_NUT*/, pps_rbsp(width, height, qp)) < 0) return WRENC_B200_EINVAL;
    return hand_out(v, out, cap);
}

// PH NAL unit + IDR_W_RADL slice NAL unit of one picture, main.rs:297-316,380-389
int64_t wrenc_b200_write_picture(int32_t qp, uint64_t picture_index, const uint8_t *slice_data, size_t len, uint8_t *out, size_t cap) {
    if (qp < -1 || qp > 63 || (len && !slice_data)) return WRENC_B200_EINVAL;
    std::vector<uint8_t> v;
    if (emit(v, 9, 19 /*PH_NUT*/, ph_rbsp(picture_index)) < 0) return WRENC_B200_EINVAL;
    std::vector<uint8_t> payload = slice_header_bits(qp);
    payload.insert(payload.end(), slice_data, slice_data + len);
    if (emit(v, 9, 7 /*IDR_W_RADL*/, payload) < 0) return WRENC_B200_EINVAL;
    return hand_out(v, out, cap);
}

// The raw RBSP / header bits (before NAL wrapping), for spec-level parse-back tests: which = 0 VPS, 1 SPS, 2 PPS, 3 PH,
// 4 slice header.
int64_t wrenc_b200_header_rbsp(int32_t which, int32_t width, int32_t height, int32_t qp, uint64_t picture_index, uint8_t *out, size_t cap) {
    if (!geometry_ok(width, height, qp)) return WRENC_B200_EINVAL;
    switch (which) {
        case 0: return hand_out(vps_rbsp(), out, cap);
        case 1: return hand_out(sps_rbsp(width, height), out, cap);
        case 2: return hand_out(pps_rbsp(width, height, qp), out, cap);
        case 3: return hand_out(ph_rbsp(picture_index), out, cap);
        case 4: return hand_out(slice_header_bits(qp), out, cap);
    }
    return WRENC_B200_EINVAL;
}
}
