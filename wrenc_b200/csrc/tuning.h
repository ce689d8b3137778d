// tuning.h — host-side derivation of every constant the search reads (product code; no oracle dependency).
// Follows BlockSplitter::new (reference src/block_splitter.rs:20-62), the header-bit heuristics of
// get_intra_pred_cost / get_chroma_intra_pred_cost (src/block_splitter.rs:187-406,594-712,775-778) and
// Quantizer::new / quantize (src/quantizer.rs:15-26,617-622,650-683).  All pow/powf calls happen HERE, on the host,
// with libm — never on the device.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

namespace wb {

struct Tuning {
    double lv_pow = 0.48592678233563835, lv_offset = 0.15150746310196822;
    float non_planar_offset = 2.2153597f, mpm_idx_offset = 1.3660221f, mpm_remainder_mult = 0.5007182f,
          mpm_remainder_offset = 2.2973304f, planar_offset = 0.9626864f, header_bits = 1.1772872f,
          chroma_header_bits = 1.309252f, qp_div = 4.4043665f, lambda_mul = 1.1282581f, cclm_pow = 0.4587651f,
          mpm_idx_pow = 0.40271285f, mpm_remainder_pow = 0.34385094f, cclm_mode_idx_offset = 2.1f, non_cclm_offset = 0.89f,
          cclm_offset = 0.53f;
    bool has_a = false;
    float a = 0.f;
    double quant_lv_pow = 0.5004010166085378, quant_qp_div = 5.218413785332902, quant_lambda_mul = 1.2709404305806742;
    long long quant_lambda_offset = 11;

    // "k=v,k=v" as main.rs:202-216 splits it; unknown keys are accepted and ignored (the reference stores them unread).
    bool parse(const char *s, std::string &err) {
        if (!s || !*s) return true;
        std::string str(s);
        size_t pos = 0;
        for (;;) {
            size_t e = str.find(',', pos);
            if (e == std::string::npos) e = str.size();
            std::string kv = str.substr(pos, e - pos);
            size_t eq = kv.find('=');
            if (eq == std::string::npos || kv.find('=', eq + 1) != std::string::npos) {
                err = "Invalid extra-params: " + str;
                return false;
            }
            std::string k = kv.substr(0, eq);
            const char *v = kv.c_str() + eq + 1;
            struct { const char *name; float *f; } fk[] = {
                {"non_planar_offset_dq_trellis", &non_planar_offset}, {"mpm_idx_offset_dq_trellis", &mpm_idx_offset},
                {"mpm_remainder_mult_dq_trellis", &mpm_remainder_mult}, {"mpm_remainder_offset_dq_trellis", &mpm_remainder_offset},
                {"planer_offset_dq_trellis", &planar_offset}, {"header_bits_dq_trellis", &header_bits},
                {"chroma_header_bits_dq_trellis", &chroma_header_bits}, {"qp_div_dq_trellis", &qp_div},
                {"lambda_mul_dq_trellis", &lambda_mul}, {"cclm_pow", &cclm_pow}, {"mpm_idx_pow", &mpm_idx_pow},
                {"mpm_remainder_pow", &mpm_remainder_pow}, {"cclm_mode_idx_offset_dq_trellis", &cclm_mode_idx_offset},
                {"non_cclm_offset_dq_trellis", &non_cclm_offset}, {"cclm_offset_dq_trellis", &cclm_offset}};
            bool hit = false;
            for (auto &f : fk)
                if (k == f.name) { *f.f = strtof(v, nullptr); hit = true; }
            if (!hit) {
                if (k == "lv_pow_dq_trellis") lv_pow = strtod(v, nullptr);
                else if (k == "lv_offset_dq_trellis") lv_offset = strtod(v, nullptr);
                else if (k == "a") { has_a = true; a = strtof(v, nullptr); }
                else if (k == "quant_lv_pow") quant_lv_pow = strtod(v, nullptr);
                else if (k == "quant_qp_div_trellis") quant_qp_div = strtod(v, nullptr);
                else if (k == "quant_lambda_mul_trellis") quant_lambda_mul = strtod(v, nullptr);
                else if (k == "quant_lambda_offset_trellis") quant_lambda_offset = strtoll(v, nullptr, 10);
            }
            if (e == str.size()) break;
            pos = e + 1;
        }
        return true;
    }
};

// Device-visible constant block (copied to global memory once per handle).
struct DevTables {
    int32_t ldq[1024];          // lambda_q * dq_table[bits]            (quantizer.rs:29-31 second term)
    int32_t lv[1024];           // lv_dq_trellis_table                  (block_splitter.rs:45-53)
    long long hdr_single[67][4];  // ((header_bits + mode_bits) * 16384.0) as i64, [luma kind][cclm kind]
    long long hdr_dual[67];       // ((header_bits / 3 + mode_bits) * 16384.0) as i64 (DUAL_TREE_LUMA)
    long long hdr_chroma[4];      // ((chroma_header_bits + cclm_bits') * 16384.0) as i64
    float lambda_rd, lambda_rd_c;
    int32_t ls;
    int32_t qp;
};

struct HostConsts {
    DevTables t;
    long long lambda_q;
    long long lv64[1024], dq64[1024];
    // returns false if a table entry does not fit the device's 32-bit storage (absurd tuning values)
    bool init(int qp, const Tuning &u, std::string &err) {
        memset(&t, 0, sizeof(t));
        t.qp = qp;
        lambda_q = (long long)(std::pow(2.0, (double)qp / u.quant_qp_div) * u.quant_lambda_mul) + u.quant_lambda_offset;
        for (int i = 0; i < 1024; i++) {
            lv64[i] = (long long)(std::pow((double)i + u.lv_offset, u.lv_pow) * 16384.0);
            dq64[i] = (long long)std::pow((double)(i * 16384), u.quant_lv_pow);
            long long l = lambda_q * dq64[i];
            // 2^26: the dependent quantisation keeps its path costs in int32 relative to the running minimum (search_kernel.cuh, trellis):
            // a step adds at most 128 * |tc - dequant| + ldq <= 2^22.2 + 2^26, the four states of a position are at most two steps
            // apart, and TR_INF = 2^28 must exceed three such steps
            if (l < 0 || l > (1ll << 26) || lv64[i] < 0 || lv64[i] > (1ll << 30)) {
                err = "tuning constants out of the supported range";
                return false;
            }
            t.ldq[i] = (int32_t)l;
            t.lv[i] = (int32_t)lv64[i];
        }
        float l = powf(2.0f, (float)qp / u.qp_div);
        t.lambda_rd = l * u.lambda_mul;
        t.lambda_rd_c = u.has_a ? l * u.a : t.lambda_rd;
        static const int kLevelScale[6] = {40, 45, 51, 57, 64, 72};
        t.ls = (16 * kLevelScale[(qp + 1) % 6]) << ((qp + 1) / 6);
        auto cclm_bits = [&](int ck) -> float {
            return ck == 0 ? u.non_cclm_offset : u.cclm_offset + powf((float)(ck - 1) + u.cclm_mode_idx_offset, u.cclm_pow);
        };
        for (int lk = 0; lk < 67; lk++) {
            float luma;
            if (lk == 0) luma = u.planar_offset;
            else if (lk <= 5) luma = u.non_planar_offset + powf((float)(lk - 1) + u.mpm_idx_offset, u.mpm_idx_pow);
            else luma = u.non_planar_offset + u.mpm_remainder_mult * powf((float)(lk - 6) + u.mpm_remainder_offset, u.mpm_remainder_pow);
            for (int ck = 0; ck < 4; ck++) {
                float mode_bits = luma + cclm_bits(ck);
                t.hdr_single[lk][ck] = (long long)((u.header_bits + mode_bits) * 16384.0f);
            }
            float mode_bits = luma + 0.0f;
            t.hdr_dual[lk] = (long long)((u.header_bits / 3.0f + mode_bits) * 16384.0f);
        }
        for (int ck = 0; ck < 4; ck++) t.hdr_chroma[ck] = (long long)((u.chroma_header_bits + cclm_bits(ck)) * 16384.0f);
        return true;
    }
};

}  // namespace wb
