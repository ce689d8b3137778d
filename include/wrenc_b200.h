/* wrenc_b200.h — C ABI of the B200-native all-intra RD search (+ slice_data coder) for hjmkt/wrenc.
 *
 * Drop-in boundary.  The reference has no plugin/FFI interface; the narrowest seam is the raster CTU loop of
 * SliceEncoder::encode (reference src/slice_encoder.rs:352-414), which for every CTU calls CtuEncoder::encode
 * (src/ctu_encoder.rs:33-202) = BlockSplitter::split_ct (src/block_splitter.rs:782-1154, the RD search) followed by the
 * syntax/CABAC writer (src/ctu_encoder.rs:227-2269).  A picture's slice_data() depends only on (Y,Cb,Cr, W, H, QP,
 * max_split_depth, tuning constants), so this library replaces that loop picture by picture:
 *
 *   wrenc_b200_create   <->  BlockSplitter::new + Quantizer::new  (src/block_splitter.rs:20-62, src/quantizer.rs:15-26),
 *                            Args --qp/--max-split-depth/--extra-params (src/main.rs:85-115,193-216)
 *   wrenc_b200_submit   <->  Picture::new + plane read + init_ctus (src/main.rs:318-361, src/picture.rs:170-195)
 *   wrenc_b200_receive  <->  SliceEncoder::encode's CTU loop result (src/slice_encoder.rs:352-419) and
 *                            Picture::get_reconst_pixels (src/picture.rs:248, --reconst dump src/main.rs:391-401)
 *   wrenc_b200_decisions<->  the decided CodingTree / TransformUnit.quantized_transformed_coeffs of every CTU
 *                            (src/ctu.rs:324-367,1190-1330,1794-1960) in flat form — the phase-1 / parity view
 *
 * Rules: plain pointers and sizes only; every function returns 0 on success or a negative error code (message via
 * wrenc_b200_last_error); nothing throws across the boundary; a handle is not thread-safe; output pointers are owned by the
 * handle and stay valid until the next receive/flush/destroy on that handle; inputs are copied before submit returns.
 * There is NO CPU fallback: create fails with WRENC_B200_ENODEV when no CUDA device is usable.
 */
#ifndef WRENC_B200_H
#define WRENC_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct wrenc_b200 wrenc_b200;

enum {
    WRENC_B200_OK = 0,
    WRENC_B200_EINVAL = -1,  /* bad argument / malformed extra_params (reference: main.rs:206-214 prints and exits) */
    WRENC_B200_ENODEV = -2,  /* no CUDA device / not sm_100 capable */
    WRENC_B200_ECUDA = -3,   /* CUDA runtime error */
    WRENC_B200_EAGAIN = -4,  /* receive with nothing submitted */
    WRENC_B200_EFULL = -5,   /* submit while every batch slot holds pictures that have not been received (call receive first) */
    WRENC_B200_EOVERFLOW = -6 /* slice_data of a picture did not fit its output buffer (twice the raw picture); the picture is still
                                 consumed: pic_idx, reconstruction and decisions are returned, *len = 0, and the handle stays usable */
};

typedef struct {
    int32_t width, height;        /* luma size; multiples of 32 (reference README.md:40) */
    int32_t qp;                   /* slice QP == --qp (reference default 26 when absent) */
    int32_t max_split_depth;      /* 0..3, --max-split-depth (default 3: CUs 32,16,8,4) */
    int32_t device;               /* CUDA device ordinal (one handle per GPU; shard picture ranges across handles) */
    int32_t pictures_in_flight;   /* pictures per batch = per search-kernel launch (CTU wavefronts of all of them interleave); two batches
                                     are in flight: one is searched while the next is filled and the previous is coded / copied back */
    int32_t want_recon;           /* copy reconstructed planes back (--reconst) */
    int32_t want_decisions;       /* copy per-CTU records + quantised levels back (parity / phase-1 consumers) */
    int32_t want_slice_data;      /* CABAC-code the decided pictures on the device and return slice_data() bytes */
    const char *extra_params;     /* "k=v,k=v" or NULL: keys of reference --extra-params */
} wrenc_b200_config;

/* One record per CTU, raster order (88 bytes, little endian). */
typedef struct {
    uint32_t split_mask;     /* bit0: 32x32 split; bits1..4: 16x16 #i split (z-order); bits5..20: 8x8 #(4i+j) split */
    uint8_t luma_mode[64];   /* intra luma mode of the CU covering each 4x4 block (8x8 raster grid) */
    uint8_t chroma_mode[16]; /* chroma prediction mode used (luma-derived value or 81=LT_CCLM,82=L_CCLM,83=T_CCLM) per 8x8 */
    float cost;              /* RD cost split_ct returned for the CTU root */
} wrenc_b200_ctu_record;

int wrenc_b200_create(const wrenc_b200_config *cfg, wrenc_b200 **out);
void wrenc_b200_destroy(wrenc_b200 *h);
const char *wrenc_b200_last_error(const wrenc_b200 *h); /* h may be NULL: error of the last failed create */

/* Host planes, tightly packed I420 as main.rs:320-349 reads them.  The planes are copied before submit returns and the
 * host-to-device copy is started at once; when the batch is complete (pictures_in_flight pictures) it is launched without
 * waiting for anything, so the caller can keep submitting the next batch while this one runs (streaming: the reference's
 * per-picture loop main.rs:294-402 becomes "read ahead and submit; receive and append" with up to 2 x pictures_in_flight
 * pictures between the reader and the writer).  May allocate (first use of a slot). */
int wrenc_b200_submit(wrenc_b200 *h, uint64_t pic_idx, const uint8_t *y, const uint8_t *cb, const uint8_t *cr);
/* The same for planes that already lie in page-locked host memory (cudaHostAlloc / cudaHostRegister, e.g. the buffer the YUV
 * reader fills): no staging copy, the planes are read by the copy engine asynchronously and must stay valid and unchanged
 * until the picture has been received.  WRENC_B200_EINVAL if a plane is not page-locked. */
int wrenc_b200_submit_pinned(wrenc_b200 *h, uint64_t pic_idx, const uint8_t *y, const uint8_t *cb, const uint8_t *cr);
/* Page-locked host memory for the caller's frame reader (what submit_pinned wants): cudaHostAlloc / cudaFreeHost without making
 * the caller link the CUDA runtime.  alloc returns NULL on failure. */
void *wrenc_b200_alloc_pinned(size_t bytes);
void wrenc_b200_free_pinned(void *p);
/* Strictly in submit order.  Launches the pending batch if it has not run yet, then blocks until THIS picture's outputs have
 * landed in host memory (pictures of one batch become ready together when its coder kernel ends; their bytes are then copied
 * back picture by picture, each with its own event).  slice_data/len: CABAC-coded slice_data() bytes of the picture (byte
 * aligned, ends with rbsp stop bit + alignment zeros).  rec_*: reconstructed planes (NULL unless want_recon).  Any out pointer
 * may be NULL.  The pointers stay valid until 2 x pictures_in_flight further pictures have been submitted, or destroy. */
int wrenc_b200_receive(wrenc_b200 *h, uint64_t *pic_idx, const uint8_t **slice_data, size_t *len, const uint8_t **rec_y,
                       const uint8_t **rec_cb, const uint8_t **rec_cr);
/* Decisions of the picture returned by the LAST receive (want_decisions must be set): records[(H/32)*(W/32)], and the
 * final quantised levels as int16 planes (luma W*H, chroma W/2*H/2), each TB stored at its own position. */
int wrenc_b200_decisions(wrenc_b200 *h, const wrenc_b200_ctu_record **records, const int16_t **lev_y, const int16_t **lev_cb,
                         const int16_t **lev_cr);
/* Launch a partly filled batch without waiting (receive does this implicitly; a full batch is launched by submit). */
int wrenc_b200_flush(wrenc_b200 *h);
/* Number of pictures submitted and not yet received. */
int wrenc_b200_pending(const wrenc_b200 *h);

/* Allocates the workspace and uploads the wavefront work list for batches of n_pictures (device-synchronising, blocking), so
 * that later search_resident / code_resident calls with at most that many pictures neither allocate nor block.  Without it
 * the first resident call with a new (larger) n_pictures does this work itself. */
int wrenc_b200_prepare(wrenc_b200 *h, int32_t n_pictures);
/* Device-resident entry (throughput path / bench "value"): n_pictures I420 pictures already in HBM, contiguous, each
 * width*height*3/2 bytes; outputs written to device buffers of the same geometry (rec: u8, levels: i16 per sample),
 * records: n_pictures*(H/32)*(W/32).  All pointers are DEVICE pointers on cfg.device and all are required.  Runs on
 * `stream` (a cudaStream_t, or NULL for the handle's own stream).  After wrenc_b200_prepare (or an earlier call with at least
 * as many pictures) it only enqueues work: no allocation, no synchronisation.  A handle runs ONE search at a time (the
 * searches share a per-grid scratch): do not run resident calls on different streams, or resident calls and submit/receive,
 * concurrently on one handle.  Returns the number of search-kernel launches enqueued (>=1) or <0. */
int wrenc_b200_search_resident(wrenc_b200 *h, int32_t n_pictures, const uint8_t *d_yuv, uint8_t *d_rec, int16_t *d_levels,
                               wrenc_b200_ctu_record *d_records, void *stream);
/* Phase 2 on resident data: CABAC-codes the pictures the preceding wrenc_b200_search_resident call on this handle decided
 * (same n_pictures, its d_levels / d_records) into d_out[n_pictures][out_cap] bytes and d_out_len[n_pictures] (-1 = overflow).
 * (-1 = the picture's output buffer is too small, -2 = the bin arena, which is sized from earlier batches, was too small for
 * this batch: call wrenc_b200_code_resident_retry).  Device pointers (d_levels 16-byte aligned); runs on `stream` after the
 * search; does not synchronise.  Returns kernel launches enqueued (6: non-zero map, syntax, scan, syntax of the long strings,
 * compaction, arithmetic coder) or <0.
 * Replaces CtuEncoder::encode_coding_tree .. encode_residual + BoolCoder (src/ctu_encoder.rs:227-2269, src/bool_coder.rs:136-296)
 * and the end_of_slice_one_bit / byte alignment of SliceEncoder::encode (src/slice_encoder.rs:380-388,419). */
int wrenc_b200_code_resident(wrenc_b200 *h, int32_t n_pictures, const int16_t *d_levels, const wrenc_b200_ctu_record *d_records, uint8_t *d_out,
                             size_t out_cap, int32_t *d_out_len, void *stream);
/* After a code_resident whose d_out_len reported -2: grows the bin arena to the size that call measured and codes the same
 * pictures again.  Blocks (reads the measured total back).  Returns kernel launches enqueued (3) or <0. */
int wrenc_b200_code_resident_retry(wrenc_b200 *h, int32_t n_pictures, const int16_t *d_levels, const wrenc_b200_ctu_record *d_records, uint8_t *d_out,
                                   size_t out_cap, int32_t *d_out_len, void *stream);
/* Workspace the resident entry needs for n_pictures (bytes of device memory it will allocate once and keep). */
size_t wrenc_b200_workspace_bytes(const wrenc_b200 *h, int32_t n_pictures);

/* Constants the search derives on the host with libm (known-answer checks). */
typedef struct {
    int64_t lambda_q;
    float lambda_rd, lambda_rd_chroma;
    int32_t ls;
    int64_t lv[8], dq[8];
} wrenc_b200_consts;
int wrenc_b200_get_consts(const wrenc_b200 *h, wrenc_b200_consts *out);
/* The same derivation without a handle or a GPU (pure host code), plus the header-bit tables
 * ((bits * 16384.0) as i64; src/block_splitter.rs:377-406,695-712): hdr_single[67][4] indexed [luma kind][cclm kind],
 * hdr_dual[67], hdr_chroma[4]; luma kind 0 = planar, 1..5 = MPM index 0..4, 6..66 = MPM remainder 0..60;
 * cclm kind 0 = not CCLM, 1..3 = cclm_mode_idx 0..2.  Any table pointer may be NULL. */
int wrenc_b200_derive_consts(int32_t qp, const char *extra_params, wrenc_b200_consts *out, int64_t *hdr_single, int64_t *hdr_dual,
                             int64_t *hdr_chroma);

/* Byte-stream NAL unit as write_byte_stream_nal_unit_bins writes it (src/nal.rs:210-299; pure host code, no GPU): three
 * zero bytes, the start code prefix 00 00 01 (nal.rs:218-225), the two-byte NAL unit header (forbidden_zero_bit,
 * nuh_reserved_zero_bit, nuh_layer_id u(6), nal_unit_type u(5), nuh_temporal_id_plus1 u(3); nal.rs:240-268), then the
 * byte-aligned payload with the reference's emulation prevention: 00 00 0x (x <= 3) gets a 03 after the two zeros, but the
 * scan only runs while idx + 3 < len, i.e. it never looks at the last three payload bytes (nal.rs:274-298).
 * For a picture the payload is the slice header bits (byte-aligned, slice_encoder.rs:338-340) followed by the slice_data()
 * bytes wrenc_b200_receive returns; the reference writes it with nuh_layer_id 9, IDR_W_RADL (7), temporal id 0
 * (main.rs:383-389).  Returns the number of bytes written to out, or the (negated) number needed if cap is too small
 * (nothing useful is written then), or WRENC_B200_EINVAL. */
int64_t wrenc_b200_write_nal(int32_t nuh_layer_id, int32_t nal_unit_type, int32_t nuh_temporal_id, const uint8_t *payload, size_t len,
                             uint8_t *out, size_t cap);

/* Complete .vvc byte streams (pure host code, no GPU).  The reference writes, once per sequence, the
 * VPS (nuh_layer_id 1), SPS and PPS (nuh_layer_id 9) NAL units (src/main.rs:223-260: VpsEncoder / SpsEncoder / PpsEncoder::encode,
 * src/vps_encoder.rs:29-279, src/sps_encoder.rs:29-665, src/pps_encoder.rs:29-350 incl. profile_tier_level / dpb_parameters /
 * ref_pic_list_struct, src/ptl_encoder.rs, src/gci_encoder.rs, src/dpbp_encoder.rs, src/rpl_encoder.rs) and, per picture, a PH NAL
 * unit (src/ph_encoder.rs:29-459, main.rs:297-316) followed by one IDR_W_RADL slice NAL unit whose payload is the byte-aligned
 * slice header (SliceEncoder::encode_sh, src/slice_encoder.rs:32-341) + slice_data() (main.rs:380-389).  For the reference's one
 * configuration (struct defaults of src/vps.rs, sps.rs, pps.rs, picture_header.rs, slice_header.rs) these bits depend only on
 * width, height, --qp (qp = -1: flag absent, QP 26) and the picture index (ph_pic_order_cnt_lsb = index & 15).
 * Each returns the bytes written, or the negated size needed when cap is too small, or WRENC_B200_EINVAL. */
int64_t wrenc_b200_write_parameter_sets(int32_t width, int32_t height, int32_t qp, uint8_t *out, size_t cap);
int64_t wrenc_b200_write_picture(int32_t qp, uint64_t picture_index, const uint8_t *slice_data, size_t len, uint8_t *out, size_t cap);
/* The header bits before NAL wrapping (spec-level parse-back tests): which = 0 VPS, 1 SPS, 2 PPS, 3 PH RBSP, 4 slice header. */
int64_t wrenc_b200_header_rbsp(int32_t which, int32_t width, int32_t height, int32_t qp, uint64_t picture_index, uint8_t *out, size_t cap);

/* Per-block entry points (host pointers): the block operations of the search, exposed for bit-exactness tests against
 * IntraPredictor::predict (src/intra_predictor.rs:56-144), Transformer::transform / inverse_transform
 * (src/transformer.rs:2040,2380), Quantizer::quantize / dequantize (src/quantizer.rs:519,761) and the rate walk of
 * get_intra_pred_cost (src/block_splitter.rs:415-460).  Blocks are n x n int16, row-major, `count` of them back to back.
 * predict: rec_i420 is a full width x height I420 reconstruction; (x,y,w) the luma TU; tree 0/1/2 = SINGLE/DUAL_LUMA/
 * DUAL_CHROMA; ar/bl the tree-position availability flags (src/ctu.rs:2083-2188); c the component; mode 0..66 or 81..83. */
int wrenc_b200_block_predict(wrenc_b200 *h, const uint8_t *rec_i420, int x, int y, int w, int tree, int ar, int bl, int c, int mode, uint8_t *pred);
int wrenc_b200_block_fwd_dct(wrenc_b200 *h, const int16_t *res, int log2n, int count, int16_t *coef);
int wrenc_b200_block_inv_dct(wrenc_b200 *h, const int16_t *deq, int log2n, int count, int16_t *out);
int wrenc_b200_block_quantize(wrenc_b200 *h, const int16_t *coef, int log2n, int count, int16_t *levels, int32_t *rates);
int wrenc_b200_block_dequantize(wrenc_b200 *h, const int16_t *levels, int log2n, int count, int16_t *out);

/* INT32 multiply-add issue rate of the device, measured with independent IMAD chains on every SM (the roofline
 * denominator of SURVEY.md §8d; one multiply-add = 2 integer ops). */
int wrenc_b200_measure_int32_peak(int device, double *imad_per_s);

const char *wrenc_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
