/*
 * dcta.h -- C ABI of libdcta.so: the B200 (sm_100a) kernels behind the dct-autoencoder
 * encode/decode transform path.
 *
 * The reference (theAdamColton/dct-autoencoder) is pure Python: there is no FFI to mirror.  The
 * drop-in boundary is the Python surface listed in SURVEY.md 8(b); this header is what that
 * surface binds to (dct_autoencoder_b200/_lib.py, ctypes).  Each entry point names the reference
 * lines it replaces (paths relative to the reference's dct_autoencoder/ package: FE =
 * feature_extraction_dct_autoencoder.py, UT = util.py, PN = patchnorm.py, VQ = vector_quantize.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors are dense, row-major, in the shapes written next to them; float = fp32,
 *     i64 = int64_t (the reference's index dtype), u8 = uint8_t (torch.bool storage);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it;
 *   - nothing is allocated or freed: outputs and workspaces are caller-owned;
 *   - return value: 0 = ok, <0 = DCTA_ERR_*; dcta_last_error() gives the message of the last
 *     failure on the calling thread.  No call throws, no call synchronises the device.
 */
#ifndef DCTA_H_
#define DCTA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCTA_OK 0
#define DCTA_ERR_INVALID_ARG (-1)
#define DCTA_ERR_LAUNCH (-2)
#define DCTA_ERR_UNSUPPORTED (-3)

const char* dcta_last_error(void);
/* Device time of every launch group between the two calls, measured with CUDA events recorded on the launching
 * stream after each group (no profiler, no synchronisation until dcta_profile_end).  One thread at a time.
 *   dcta_profile_end: names_out receives the group names joined by '\n' (NUL-terminated, at most names_cap bytes),
 *   ms_out up to max_n durations in launch order; returns the number of groups. */
int dcta_profile_begin(void* stream);
int dcta_profile_end(char* names_out, int names_cap, float* ms_out, int max_n);
/* ABI version of this header; bumped on any signature change. */
int dcta_abi_version(void);
/* Compute capability the library was compiled for (100 for sm_100a). */
int dcta_compiled_arch(void);

/* ------------------------------------------------------------------ basis tables ---------- */
/* The orthonormal DCT-II matrix behind UT:333-338 (torch_dct "ortho"): C_n[q, m] = s_q cos(pi (2m+1) q / 2n),
 * evaluated in double on the HOST and written, in the layout the named GEMM path reads, into caller-provided
 * HOST buffers (no allocation, no device work: upload once per (layout, n, k) and keep the device copies).
 *   DCTA_BASIS_F32        hi = float (k, n)                                  -> dcta_dct2_fwd / dcta_dct2_inv
 *   DCTA_BASIS_SPLIT_FWD  hi, lo = fp16 (k, round8(n)), C * 2^10 with row 0 stored as the constant 32;
 *                         row_scale (k) undoes the scaling                    -> dcta_dct2_fwd_tc
 *   DCTA_BASIS_SPLIT_INV  hi, lo = fp16 (n, round8(k)) = the transpose        -> dcta_dct2_inv_tc
 *   DCTA_BASIS_FOLD_FWD   hi, lo = fp16 (2, k/2, n/2): group g holds rows 2j+g over the first half of the samples;
 *                         row_scale (2, k/2)                                  -> dcta_dct2_fwd_fold[_codes]
 *   DCTA_BASIS_FOLD_INV   hi, lo = fp16 (2, n/2, round8(k/2)) = the transposes -> dcta_dct2_inv_fold
 * hi + lo carries 22 significand bits of every entry.  dcta_basis_elems: elements per plane (-1: unknown layout). */
#define DCTA_BASIS_F32 0
#define DCTA_BASIS_SPLIT_FWD 1
#define DCTA_BASIS_SPLIT_INV 2
#define DCTA_BASIS_FOLD_FWD 3
#define DCTA_BASIS_FOLD_INV 4
int64_t dcta_basis_elems(int layout, int n, int k);
int dcta_basis_init(int layout, int n, int k, void* hi_host, void* lo_host, float* row_scale_host);

/* ------------------------------------------------------------------ colour space ---------- */
/* UT:70-82 rgb_to_ipt (rgb_to_lms UT:56-60, channel_mult UT:46-47).
 * rgb, ipt: (n_img, 3, plane).  m_rgb2lms_host, m_ipt_host: 3x3 row-major fp32 HOST matrices
 * (the reference's Trgb2lms UT:40 and Mipt UT:37-39).  rgb and ipt must not overlap. */
int dcta_rgb_to_ipt(const float* rgb, float* ipt, int64_t n_img, int64_t plane,
                    const float* m_rgb2lms_host, const float* m_ipt_host, void* stream);
/* 8-bit pixels -> floats in [0, 1]: out[i] = in[i] / 255 with torch's IEEE fp32 division, bit for bit (the callers'
 * `torchvision.io.read_image(path) / 255`, decode_gif.py:22, testpipe.py:17), and back the way
 * torchvision.utils.save_image stores a float image (testpipe.py:74-75): floor(clamp(x, 0, 1) * 255 + 0.5). */
int dcta_u8_to_unit_f32(const uint8_t* in, float* out, int64_t n, void* stream);
int dcta_unit_f32_to_u8(const float* in, uint8_t* out, int64_t n, void* stream);
/* UT:85-97 ipt_to_rgb.  m_ipt_inv_host = Mipt.inverse() (UT:91), m_lms2rgb_host = Tlms2rgb (UT:41). */
int dcta_ipt_to_rgb(const float* ipt, float* rgb, int64_t n_img, int64_t plane,
                    const float* m_ipt_inv_host, const float* m_lms2rgb_host, void* stream);

/* ------------------------------------------------------------------ DCT / IDCT ------------ */
/* Truncated orthonormal 2-D DCT-II of n_planes planes (FE:140 + the crop of FE:348-362 and the
 * max_patch clip of FE:393 folded into the basis):  Y = CH[:kh] . X . CW[:kw]^T.
 *   x    (n_planes, h, w)
 *   ch   (kh, h), cw (kw, w)      basis rows (host-generated in float64, rounded to fp32)
 *   work (n_planes, h, kw)        scratch
 * Output layout:
 *   tile_p == 0 : y (n_planes, kh, kw) planes
 *   tile_p  > 0 : token-grid layout of FE:374-380, y (n_planes / channels, kh/p, kw/p, channels, p*p)
 *                 (kh, kw multiples of p; n_planes multiple of channels). */
int dcta_dct2_fwd(const float* x, const float* ch, const float* cw, float* work, float* y,
                  int64_t n_planes, int h, int w, int kh, int kw, int tile_p, int channels,
                  void* stream);
/* Truncated inverse (FE:149 with the zero padding of FE:300-304 folded into the basis):
 *   X = CH[:kh]^T . Y . CW[:kw];  y (n_planes, kh, kw), work (n_planes, h, kw), x (n_planes, h, w). */
int dcta_dct2_inv(const float* y, const float* ch, const float* cw, float* work, float* x,
                  int64_t n_planes, int h, int w, int kh, int kw, void* stream);

/* ------------------------------------------------------------------ DCT / IDCT on tensor cores */
/* Split-precision path (csrc/gemm_tc.cu): every fp32 operand is carried as two fp16 planes
 * v * 2^s = hi + lo and each contraction is three tcgen05.mma (hi*hi + hi*lo + lo*hi) accumulated in
 * fp32 in TMEM.  `half` pointers are passed as void*.  All hi/lo planes are K-major with an element
 * pitch that is a multiple of 8 and a 16-byte aligned base (TMA).
 *
 * dcta_gemm_split: D[b] (a_rows x b_rows, fp32, pitch out_ld) = alpha * row_scale[m] *
 *   sum_k A[b][m,k] * B[b][n,k] + col_bias[n];  a/b_batch_stride == 0 means the operand is shared by all b.
 *   a_lo [nullable]: NULL = the A values are exact fp16 (e.g. the +-1 LFQ codes): two MMAs per k step, no A_lo loads.
 *   row_scale, col_bias [nullable]. */
int dcta_gemm_split(const void* a_hi, const void* a_lo, int a_rows, int64_t a_ld, int64_t a_batch_stride,
                    const void* b_hi, const void* b_lo, int b_rows, int64_t b_ld, int64_t b_batch_stride,
                    int k, int64_t batch, const float* row_scale, float alpha, const float* col_bias, float* out,
                    int64_t out_ld, int64_t out_batch_stride, void* stream);
/* LFQ with projections in eval (lfq.py:136-227, has_projections): project_in GEMM + bias + sign in one kernel.
 *   a_hi/a_lo (rows, a_ld): split rows of the tokens (dcta_split_rows_rowscale) with their row_scale;
 *   w_hi/w_lo (n, w_ld): project_in.weight split; bias (n) [nullable];
 *   q_hi (rows, q_ld) fp16 = +-codebook_scale: the A operand of project_out (pass a_lo = NULL to dcta_gemm_split);
 *   sign_bits (rows, ceil(n/128), 4) uint32 -> dcta_lfq_bits_to_codes: indices (rows, c) int64, MSB first (LFQ:87, 187). */
int dcta_lfq_project_sign(const void* a_hi, const void* a_lo, int64_t rows, int64_t a_ld, const void* w_hi,
                          const void* w_lo, int n, int64_t w_ld, int k, const float* row_scale, const float* bias,
                          float codebook_scale, void* q_hi, int64_t q_ld, uint32_t* sign_bits, void* stream);
int dcta_lfq_bits_to_codes(const uint32_t* sign_bits, int64_t rows, int n, int c, int d, int64_t* codes, void* stream);
/* hi = rn16(x*scale), lo = rn16(x*scale - hi), elementwise over n values. */
int dcta_split_f32(const float* x, void* hi, void* lo, int64_t n, float scale, void* stream);
/* The tensor core accumulates in fp32 with truncation, so the producers below remove a per-plane
 * constant (the plane mean / the DC coefficient) before the split and pass it in `dc`; the GEMM
 * epilogue adds its exact contribution back (the DCT of a constant is the DC coefficient only).
 * DCTA_SUM_SCRATCH floats of scratch per plane are needed by the mean estimate. */
#define DCTA_SUM_SCRATCH 33
/* fp32 planes (n_planes, h, w) -> centred hi/lo (scale 2^8) + dc[n_planes] = mean * sqrt(h*w). */
int dcta_split_planes_centered(const float* x, void* hi, void* lo, float* dc, float* sums_scratch,
                               int64_t n_planes, int h, int w, void* stream);
/* fp32 coefficient planes (n_planes, kh, kw) -> hi/lo (n_planes, kh, ld) scale 2^4 with the DC
 * coefficient moved to dc[n_planes] = Y[0,0] / sqrt(out_h*out_w). */
int dcta_split_coef_planes(const float* y, void* hi, void* lo, float* dc, int64_t n_planes, int kh,
                           int kw, int64_t ld, int out_h, int out_w, void* stream);
/* UT:70-82 rgb_to_ipt producing the forward GEMM's operand planes directly: (ipt - mean)*2^8 as
 * hi/lo (n_img, 3, h, w) and dc (n_img*3); h*w % 4 == 0; sums_scratch (n_img*3*DCTA_SUM_SCRATCH). */
int dcta_rgb_to_ipt_split(const float* rgb, void* ipt_hi, void* ipt_lo, float* dc, float* sums_scratch,
                          int64_t n_img, int h, int w, const float* m_rgb2lms_host,
                          const float* m_ipt_host, void* stream);
/* FE:635-653 un-patchify producing the inverse GEMM's operand planes: y*2^4 as hi/lo
 * (n_img, channels, rows, ld), columns >= cols zero-filled, DC moved to dc (n_img*channels) scaled
 * by 1/sqrt(out_h*out_w).  img_sel (n_img) i32 [nullable = identity] picks images of slot_map. */
int dcta_unpatchify_split(const float* patches, const int32_t* slot_map, const int32_t* img_sel,
                          int64_t n_img, int channels_n, int th, int tw, int p, int rows, int cols,
                          int64_t ld, int out_h, int out_w, void* y_hi, void* y_lo, float* dc, void* stream);
/* Forward truncated DCT (same contract as dcta_dct2_fwd) from centred split planes.
 *   x_hi/lo (n_planes, h, w) scale 2^8, dc (n_planes) [nullable];  bw = CW'[:kw] (kw, w),
 *   bh = CH'[:kh] (kh, ld_h): basis * 2^10 with row 0 stored as the constant 32 (exact) -- rs_w / rs_h
 *   (kw / kh floats) undo those scales;  work_hi/lo (n_planes, kw, ld_h). */
int dcta_dct2_fwd_tc(const void* x_hi, const void* x_lo, const float* dc, const void* bw_hi,
                     const void* bw_lo, const float* rs_w, const void* bh_hi, const void* bh_lo,
                     const float* rs_h, void* work_hi, void* work_lo, float* y, int64_t n_planes, int h,
                     int w, int kh, int kw, int64_t ld_h, int tile_p, int channels, void* stream);
/* Inverse truncated DCT (same contract as dcta_dct2_inv) from split coefficient planes.
 *   y_hi/lo (n_planes, kh, ld_kw) scale 2^4, dc (n_planes) [nullable];  bwt = (CW' * 2^10)^T
 *   (w, ld_kw), bht = (CH' * 2^10)^T (h, ld_kh); work_hi/lo (n_planes, w, ld_kh);  x (n_planes, h, w). */
int dcta_dct2_inv_tc(const void* y_hi, const void* y_lo, const float* dc, const void* bwt_hi,
                     const void* bwt_lo, const void* bht_hi, const void* bht_lo, void* work_hi,
                     void* work_lo, float* x, int64_t n_planes, int h, int w, int kh, int kw,
                     int64_t ld_kh, int64_t ld_kw, void* stream);

/* ------------------------------------------------------------------ folded DCT / IDCT ----- */
/* The DCT-II basis is (anti)symmetric about the centre of the signal, so the 2-D transform of an
 * h x w plane (UT:333-338, called at FE:140 / FE:149) splits into four independent transforms of
 * (h/2, w/2) "quadrants" xq[b][a] = the four sign combinations of the mirrored samples
 * x[h',w'], x[h',w-1-w'], x[h-1-h',w'], x[h-1-h',w-1-w']:   Y[2i+a, 2j+b] = CH[2i+a,:h/2] . xq[b][a] . CW[2j+b,:w/2]^T
 * -- half the multiply-adds.  Requirements: h, w multiples of 16; kh, kw even.
 * dcta_fold_supported: 1 if these sizes can take the folded path (else use the dcta_*_tc entry points). */
int dcta_fold_supported(int h, int w, int kh, int kw);
/* UT:70-82 rgb_to_ipt + centring + fold + fp16 hi/lo split: rgb (n_img, 3, h, w) ->
 * xq_hi/lo (2, 2, n_img*3, h/2, w/2) [b][a][plane] scale 2^6;  dc, sums_scratch as dcta_rgb_to_ipt_split. */
int dcta_rgb_to_ipt_fold(const float* rgb, void* xq_hi, void* xq_lo, float* dc, float* sums_scratch,
                         int64_t n_img, int h, int w, const float* m_rgb2lms_host, const float* m_ipt_host,
                         void* stream);
/* The same for 8-bit pixels: rgb (n_img, 3, h, w) uint8, read as u8 / 255 exactly as torch's IEEE fp32 division
 * does (the callers' `torchvision.io.read_image(path) / 255`, decode_gif.py:22, testpipe.py:17).  4x fewer bytes
 * over PCIe for host-resident images. */
int dcta_rgb_u8_to_ipt_fold(const uint8_t* rgb, void* xq_hi, void* xq_lo, float* dc, float* sums_scratch,
                            int64_t n_img, int h, int w, const float* m_rgb2lms_host, const float* m_ipt_host,
                            void* stream);
/* The same for plain fp32 planes x (n_planes, h, w). */
int dcta_fold_planes(const float* x, void* xq_hi, void* xq_lo, float* dc, float* sums_scratch,
                     int64_t n_planes, int h, int w, void* stream);
/* Forward truncated DCT from the quadrants; output contract of dcta_dct2_fwd (planes or token grid).
 *   bw_hi/lo (2, kw/2, w/2): CW[2j+b, w'] * 2^10, group b (row 0 of group 0 stored as the constant 32);
 *   rs_w (2, kw/2) the factors undoing those scales;  bh_hi/lo (2, kh/2, h/2), rs_h likewise (group a);
 *   work_hi/lo (2, n_planes, kw, h/2) scratch;
 *   maxabs [nullable, token grid only] (n_planes/channels, kh/p, kw/p, channels): amax|tile| of every token
 *   (FE:409), reduced in the GEMM epilogue (input of dcta_sort_tokens_maxabs). */
int dcta_dct2_fwd_fold(const void* xq_hi, const void* xq_lo, const float* dc, const void* bw_hi,
                       const void* bw_lo, const float* rs_w, const void* bh_hi, const void* bh_lo,
                       const float* rs_h, void* work_hi, void* work_lo, float* y, float* maxabs,
                       int64_t n_planes, int h, int w, int kh, int kw, int tile_p, int channels, void* stream);
/* The forward transform straight to LFQ code words, for a projection-free LFQ with one codebook per patch row
 * (c == d == tile_p) on frozen PatchNorm tables: the pass-2 epilogue forms the sign bit of
 * clamp((Y - median) / (b*sqrt2 + eps)) (PN:157-165, LFQ:175-187) of every coefficient and never writes the token grid.
 *   code_grid (n_planes/channels, kh/p, kw/p, channels, p) int32: code word of patch row r of every token, in
 *   token-grid order (dcta_pack_codes_grid gathers them into sorted, packed order);  maxabs as above;
 *   median, b (channels, H, W, p*p);  tame_scratch: one device int32 (the "every b is a tame divisor" flag);
 *   tame_known != 0: tame_scratch already holds the flag of these tables (set by an earlier call), skip the check. */
int dcta_fold_codes_supported(int h, int w, int kh, int kw, int tile_p);   /* 1 if the next entry point takes these sizes */
int dcta_dct2_fwd_fold_codes(const void* xq_hi, const void* xq_lo, const float* dc, const void* bw_hi,
                             const void* bw_lo, const float* rs_w, const void* bh_hi, const void* bh_lo,
                             const float* rs_h, void* work_hi, void* work_lo, float* maxabs, int32_t* code_grid,
                             const float* median, const float* b, int H, int W, float eps, float lo, float hi,
                             int32_t* tame_scratch, int tame_known, int64_t n_planes, int h, int w, int kh, int kw,
                             int tile_p, int channels, void* stream);
/* FE:635-653 un-patchify into folded coefficient quadrants yq_hi/lo (2, 2, n_img*channels, rows/2, ldq)
 * [b][a][plane][i][j] = Y[2i+a, 2j+b] * 2^4, ldq = round8(cols/2); DC moved to dc as in dcta_unpatchify_split. */
int dcta_unpatchify_fold(const float* patches, const int32_t* slot_map, const int32_t* img_sel, int64_t n_img,
                         int channels_n, int th, int tw, int p, int rows, int cols, int out_h, int out_w,
                         void* yq_hi, void* yq_lo, float* dc, void* stream);
/* The same from NORMALISED patches, with PatchNorm.inverse_norm (patchnorm.py:167-177: x * (b*sqrt2 + eps) + median) applied
 * on the way: equals dcta_unpatchify_fold(dcta_patchnorm_apply(inverse)) bit for bit, the de-normalised patches are never
 * written.  median, b (channels_n, H, W, p*p); th <= H, tw <= W. */
int dcta_unpatchify_denorm_fold(const float* patches, const int32_t* slot_map, const int32_t* img_sel, int64_t n_img,
                                int channels_n, int th, int tw, int p, int rows, int cols, int out_h, int out_w,
                                const float* median, const float* b, int H, int W, float eps, void* yq_hi, void* yq_lo,
                                float* dc, void* stream);
/* The same from LFQ codes (contract of dcta_decode_codes_split).
 *   tab_scratch [nullable]: 2 * channels*H*p*ceil(W*p/8)*8 device uint32.  When given (and c == d == p <= 16) the two
 *   values a de-quantised coefficient can take at every position are tabulated once per call instead of once per CTA. */
int dcta_decode_codes_fold(const int64_t* codes, const int32_t* slot_map, const int32_t* img_sel,
                           int64_t n_img, int channels_n, int th, int tw, int p, int rows, int cols, int out_h,
                           int out_w, const float* median, const float* b, int H, int W, float eps, int c,
                           int d, float scale, void* yq_hi, void* yq_lo, float* dc, void* tab_scratch, void* stream);
/* Decode from LFQ codes INSIDE inverse pass 1 (FE:607-656 revert_patching + PN:167-177 inverse_norm + LFQ:105-134
 * indices_to_codes + the inverse DCT of FE:289-310): the data operand of pass 1 is generated in shared memory from one
 * sign bit per coefficient and a table of the two values a de-quantised, de-normalised coefficient can take, so the
 * coefficient planes are never written.  Needs one codebook per patch row (c == d == p, p even in 8..16).
 *   dcta_decode_gen_tables: the two-value table of (median, b, eps, scale) for kh x kw coefficients,
 *     dcta_decode_gen_tables_bytes(...) bytes; depends on the PatchNorm tables only, so it can be kept across calls.
 *   dcta_decode_codes_inv_fold: arguments as dcta_decode_codes_fold + dcta_dct2_inv_fold (kh = rows, kw = cols);
 *     tab: that table;  scratch: dcta_decode_codes_inv_fold_scratch_bytes(...) bytes, 256-byte aligned.
 * Results are bit-identical to dcta_decode_codes_fold followed by dcta_dct2_inv_fold. */
int64_t dcta_decode_gen_tables_bytes(int channels_n, int kh, int kw);
int dcta_decode_gen_tables(const float* median, const float* b, int channels_n, int H, int W, float eps, int p,
                           int kh, int kw, float scale, void* tab, void* stream);
int64_t dcta_decode_codes_inv_fold_scratch_bytes(int64_t n_img, int channels_n, int kh, int kw);
int dcta_decode_codes_inv_fold_supported(int h, int w, int kh, int kw, int p, int c, int d);
int dcta_decode_codes_inv_fold(const int64_t* codes, const int32_t* slot_map, const int32_t* img_sel, int64_t n_img,
                               int channels_n, int th, int tw, int p, int kh, int kw, int h, int w,
                               const float* median, const float* b, int H, int W, float eps, int c, int d,
                               float scale, const void* bwt_hi, const void* bwt_lo, const void* bht_hi,
                               const void* bht_lo, void* work_hi, void* work_lo, float* z, float* dc,
                               const void* tab, void* scratch, void* stream);
/* The same straight from the code grid dcta_dct2_fwd_fold_codes wrote ((n_img, kh/p, kw/p, channels, p) int32), for a
 * round trip that kept every token: no slot map and no gather through the packed codes.  Same results. */
int dcta_decode_grid_inv_fold(const int32_t* code_grid, int64_t n_img, int channels_n, int p, int kh, int kw, int h,
                              int w, const float* median, const float* b, int H, int W, float eps, float scale,
                              const void* bwt_hi, const void* bwt_lo, const void* bht_hi, const void* bht_lo,
                              void* work_hi, void* work_lo, float* z, float* dc, const void* tab, void* scratch,
                              void* stream);
/* fp32 coefficient planes y (n_planes, kh, kw) -> folded quadrants. */
int dcta_fold_coef_planes(const float* y, void* yq_hi, void* yq_lo, float* dc, int64_t n_planes, int kh,
                          int kw, int out_h, int out_w, void* stream);
/* Inverse truncated DCT of the quadrants: z (4, n_planes, h/2, w/2) fp32, index b*2+a.
 *   bwt_hi/lo (2, w/2, ldq): (CW[2j+b, w'] * 2^10)^T, bht_hi/lo (2, h/2, ldi), ldi = round8(kh/2);
 *   work_hi/lo (2, 2, n_planes, w/2, ldi) scratch. */
int dcta_dct2_inv_fold(const void* yq_hi, const void* yq_lo, const void* bwt_hi, const void* bwt_lo,
                       const void* bht_hi, const void* bht_lo, void* work_hi, void* work_lo, float* z,
                       int64_t n_planes, int h, int w, int kh, int kw, void* stream);
/* Final butterfly (+ the DC constant dc[plane], nullable) fused with UT:85-97 ipt_to_rgb: rgb (n_img, 3, h, w). */
int dcta_unfold_ipt_to_rgb(const float* z, const float* dc, float* rgb, int64_t n_img, int h, int w,
                           const float* m_ipt_inv_host, const float* m_lms2rgb_host, void* stream);
/* The same with 8-bit output, rgb (n_img, 3, h, w) uint8 = floor(clamp(x, 0, 1) * 255 + 0.5): the quantisation of
 * torchvision.utils.save_image, which is what the reference's callers apply to a reconstruction (testpipe.py:74-75). */
int dcta_unfold_ipt_to_rgb_u8(const float* z, const float* dc, uint8_t* rgb, int64_t n_img, int h, int w,
                              const float* m_ipt_inv_host, const float* m_lms2rgb_host, void* stream);
/* Final butterfly alone: x (n_planes, h, w). */
int dcta_unfold_planes(const float* z, const float* dc, float* x, int64_t n_planes, int h, int w, void* stream);

/* FE:374-380 rearrange "c (h p1) (w p2) -> (h w) c (p1 p2)" with the max_patch clip of FE:393-394,
 * for coefficient planes produced by a caller-supplied transform:
 *   planes (n_img, channels, rows, cols) -> tiles (n_img, th, tw, channels, p*p); th*p<=rows, tw*p<=cols. */
int dcta_patchify(const float* planes, float* tiles, int64_t n_img, int channels_n, int rows,
                  int cols, int th, int tw, int p, void* stream);

/* ------------------------------------------------------------------ select / pack --------- */
/* FE:403-416 importance scores of every token of the token grid, in the reference's pre-sort
 * flat order (tile-major, channel-minor):  0.1*max|tile| - (th+tw)/channel_importance[c].
 *   tiles (n_img, th, tw, channels, z)  ->  scores (n_img, th*tw*channels) */
int dcta_tile_scores(const float* tiles, float* scores, int64_t n_img, int th, int tw, int channels,
                     int z, float mag_weight, const float* channel_importances_host, void* stream);
/* FE:418 per-image descending sort of the scores (ties: ascending flat index).
 *   scores (n_img, n_tok) -> order (n_img, n_tok) int32.  n_tok <= 16384. */
int dcta_sort_tokens(const float* scores, int32_t* order, int64_t n_img, int n_tok, void* stream);
/* The same from per-token amax|tile| (maxabs (n_img, th, tw, channels)): the scores of FE:409-416 are
 * formed on the fly with the arithmetic of dcta_tile_scores; scores [nullable] receives them. */
int dcta_sort_tokens_maxabs(const float* maxabs, float* scores, int32_t* order, int64_t n_img, int th, int tw,
                            int channels, float mag_weight, const float* channel_importances_host, void* stream);

/* One packed row segment: k tokens of image `img` placed at slots [offset, offset+k) of row `row`
 * with batched_image_id `image_id` (FE:455-513 next-fit result, computed on the host). */
typedef struct {
    int32_t row, offset, k, image_id;
    int64_t img; /* tile mode: image index into tiles/order; list mode: index into the ptr tables */
} dcta_segment;

/* FE:437-452 gather by sorted index + FE:516-605 / UT:149-164 pad-and-stack, in one pass.
 * Every element of every output is written exactly once (padding: zeros, key_pad_mask = 1).
 *   segs (n_seg) device array sorted by (row, offset); row_seg_start (n_rows + 1) device
 *   tiles (n_img, th*tw*channels, z), order (n_img, th*tw*channels)
 *   patches (n_rows, s, z), positions (n_rows, s, 2) i64, channels_out (n_rows, s) i64,
 *   image_ids (n_rows, s) i64 [nullable], key_pad_mask (n_rows, s) u8 [nullable] */
int dcta_pack_tiles(const float* tiles, const int32_t* order, const dcta_segment* segs,
                    const int32_t* row_seg_start, int n_rows, int s, int th, int tw, int channels,
                    int z, float* patches, int64_t* positions, int64_t* channels_out,
                    int64_t* image_ids, uint8_t* key_pad_mask, void* stream);
/* The bookkeeping of dcta_pack_tiles without the copy: positions / channels_out / image_ids / key_pad_mask as above and
 *   src_index (n_rows, s) int32 = the row of the token grid tiles (n_img * th*tw*channels, z) each slot would have read
 *   (-1 for padding slots).  A consumer that streams the rows anyway gathers them itself (dcta_split_rows_patchnorm). */
int dcta_pack_tiles_index(const int32_t* order, const dcta_segment* segs, const int32_t* row_seg_start, int n_rows, int s,
                          int th, int tw, int channels, int64_t n_img, int64_t* positions, int64_t* channels_out,
                          int64_t* image_ids, uint8_t* key_pad_mask, int32_t* src_index, void* stream);
/* Same packing for per-image token lists produced earlier by preprocess (FE:180-287 iter_batches
 * path): src_* are device arrays of n_src device pointers, indexed by seg.img. */
int dcta_pack_lists(const float* const* src_patches, const int64_t* const* src_positions,
                    const int64_t* const* src_channels, const dcta_segment* segs,
                    const int32_t* row_seg_start, int n_rows, int s, int z, float* patches,
                    int64_t* positions, int64_t* channels_out, int64_t* image_ids,
                    uint8_t* key_pad_mask, void* stream);

/* ------------------------------------------------------------------ PatchNorm ------------- */
/* PN:157-165 (inverse == 0):  clamp((x - median[c,h,w]) / (b[c,h,w]*sqrt(2) + eps), lo, hi)
 * PN:167-177 (inverse != 0):  x * (b[c,h,w]*sqrt(2) + eps) + median[c,h,w]
 *   x, out (n_tok, z); channels (n_tok) i64; positions (n_tok, 2) i64; median, b (C, H, W, z).
 * Padding tokens are transformed with the statistics at (0,0,0) exactly as the reference does. */
int dcta_patchnorm_apply(const float* x, const int64_t* channels, const int64_t* positions,
                         const float* median, const float* b, float* out, int64_t n_tok, int z,
                         int C, int H, int W, float eps, float lo, float hi, int inverse,
                         void* stream);

/* PN:101-150 statistic fitting, split where a multi-GPU sum can be inserted.
 * Step 1 -- per-position token lists (deterministic: sorted by flat token index):
 *   counts (n_pos) i32 out, offsets (n_pos + 1) i32 out, cursor (n_pos) i32 scratch,
 *   list (2 * n_tok) i32: first half = sorted lists (out), second half scratch; n_pos = C*H*W.
 *   key_pad_mask (n_tok) u8: 1 = padding, skipped (PN:105-108). */
int dcta_patchnorm_build_lists(const int64_t* channels, const int64_t* positions,
                               const uint8_t* key_pad_mask, int64_t n_tok, int C, int H, int W,
                               int32_t* counts, int32_t* offsets, int32_t* cursor, int32_t* list,
                               void* stream);
/* Step 2 -- PN:123-130 per-position batch median (lower middle, as torch.median) and PN:112-119
 * batch_n.  Writes the reduce-ready buffer  packed = [batch_n (n_pos) | batch_median*batch_n (n_pos*z)]. */
int dcta_patchnorm_batch_median(const float* x, const int32_t* offsets, const int32_t* list,
                                int n_pos, int z, float* packed, void* stream);
/* Step 3 -- PN:135-138 running-median update from the (possibly all-reduced) packed buffer:
 *   median <- (median*n + sum(batch_median*batch_n)) / clamp(n + batch_n, 1). */
int dcta_patchnorm_update_median(float* median, const float* n, const float* packed, int n_pos,
                                 int z, void* stream);
/* Step 4 -- PN:140-143: abs_dev (n_pos*z) = sum over the position's tokens |x - median_new|. */
int dcta_patchnorm_abs_dev(const float* x, const int32_t* offsets, const int32_t* list,
                           const float* median, int n_pos, int z, float* abs_dev, void* stream);
/* Step 5 -- PN:144-150 from the (possibly all-reduced) abs_dev and packed batch_n:
 *   b <- (b*n + (abs_dev/clamp(batch_n,1))*batch_n) / clamp(n + batch_n, 1);  n <- n + batch_n. */
int dcta_patchnorm_update_b(float* b, float* n, const float* packed, const float* abs_dev,
                            int n_pos, int z, void* stream);
/* PN:153-155: out = x with padding tokens zeroed. */
int dcta_zero_padding(const float* x, const uint8_t* key_pad_mask, float* out, int64_t n_tok, int z,
                      void* stream);

/* ------------------------------------------------------------------ LFQ ------------------- */
/* lfq.py:168-187 (eval): q = where(x > 0, +scale, -scale); indices[t, c] = sum_i (x>0) << (d-1-i).
 *   x, q (n_tok, c*d) [q nullable]; indices (n_tok, c) i64.  d <= 62. */
int dcta_lfq_quantize(const float* x, float* q, int64_t* indices, int64_t n_tok, int c, int d,
                      float scale, void* stream);
/* lfq.py:105-134: codes (n_tok, c*d) = bit ? +scale : -scale, MSB first. */
int dcta_lfq_indices_to_codes(const int64_t* indices, float* codes, int64_t n_tok, int c, int d,
                              float scale, void* stream);
/* lfq.py:195-200 masked commitment loss:  sum over VALID tokens of (x - q)^2, divided by
 * (n_valid * c * d).  mask (n_tok) u8, 1 = valid.  result: one float (device).
 * scratch2: DCTA_REDUCE_SCRATCH floats (deterministic two-stage reduction). */
#define DCTA_REDUCE_SCRATCH 2048
int dcta_lfq_commit_loss(const float* x, const uint8_t* mask, float* result, float* scratch2,
                         int64_t n_tok, int cd, float scale, void* stream);
/* lfq.py:191: distance (n_tok, c, 2^d) = -2 * <x[t, c, :], codebook[j, :]>  (d <= 16). */
int dcta_lfq_distance(const float* x, float* distance, int64_t n_tok, int c, int d, float scale,
                      void* stream);
/* UT:355-387 compute_entropy_loss on a dense affinity (n_tok, c, n_codes), mask (n_tok) u8 1=valid.
 * scratch: (c * n_codes + 2) floats, zeroed by the call.  result: one float (device). */
int dcta_entropy_loss(const float* affinity, const uint8_t* mask, float* scratch, float* result,
                      int64_t n_tok, int c, int n_codes, float temperature, float eps, void* stream);
/* UT:391-410 calculate_perplexity: counts (codebook_size) i64 scratch, result one float. */
int dcta_perplexity(const int64_t* codes, int64_t n, int codebook_size, int64_t null_index,
                    int64_t* counts, float* result, void* stream);

/* ------------------------------------------------------------------ code wire format ------ */
/* Compact replacement of DP:54-87 to_dict's per-token {c, h, w, data} dicts (consumer:
 * prepare_autoregressive_dataset.py:51-66).  One record of rec = 2 + ceil(c*d/8) bytes per slot of the
 * (n_rows, s) batch: little-endian u16 channel<<12 | h<<6 | w, then the token's c code words of d bits each,
 * most significant bit first, concatenated (last byte zero-padded).  Needs channel < 16 and h, w < 64.
 * counts [nullable]: (n_rows, s) int32, zeroed by the call; counts[r, i] = number of non-padding tokens of
 * the i-th image of row r (tokens of one image are adjacent, FE:516-605), which is all the host needs to
 * cut out[r] into per-image byte strings. */
int dcta_wire_pack(const int64_t* codes, const int64_t* positions, const int64_t* channels,
                   const int64_t* image_ids, const uint8_t* key_pad_mask, int64_t n_rows, int s, int c, int d,
                   int rec, uint8_t* out, int32_t* counts, void* stream);
/* Inverse (DP:90-122 from_dict): n_tok records -> codes (n_tok, c) i64, positions (n_tok, 2) i64 [h, w],
 * channels (n_tok) i64.  `in` must be 4-byte aligned. */
int dcta_wire_unpack(const uint8_t* in, int64_t n_tok, int c, int d, int rec, int64_t* codes,
                     int64_t* positions, int64_t* channels, void* stream);

/* ------------------------------------------------------------------ fused PatchNorm + LFQ -- */
/* Encode to codes without materialising the gathered / normalised / quantised patches:
 * dcta_pack_tiles' gather (FE:437-452, FE:516-605) + PN:157-165 + lfq.py:168-187 in registers.
 * Requires a projection-free LFQ (c*d == z <= 256) and frozen statistics.  Padding slots are
 * quantised as the reference does (zeros normalised with the statistics at (0,0,0)).
 * Outputs are bit-identical to dcta_pack_tiles -> dcta_patchnorm_apply -> dcta_lfq_quantize.
 * tame_scratch [nullable]: one device int32 of scratch.  When given, the call first checks on the device that
 * every b is finite and in [0, 1e18]; if so the kernel reads b only where |x - median| < 1e-20 (the bit is
 * the sign of x - median everywhere else). */
int dcta_pack_codes_lfq(const float* tiles, const int32_t* order, const dcta_segment* segs,
                        const int32_t* row_seg_start, int n_rows, int s, int th, int tw, int channels,
                        int z, const float* median, const float* b, int H, int W, float eps, float lo,
                        float hi, int c, int d, float scale, int32_t* tame_scratch, int64_t* codes, int64_t* positions,
                        int64_t* channels_out, int64_t* image_ids, uint8_t* key_pad_mask, void* stream);
/* FE:437-452 / 516-605 gather of those code words: codes (n_rows, s, c) i64 + the bookkeeping outputs of
 * dcta_pack_codes_lfq.  Padding slots get the code words of an all-zero token at position (0,0,0), as in the
 * reference.  pad_scratch: c device int64. */
int dcta_pack_codes_grid(const int32_t* code_grid, const int32_t* order, const dcta_segment* segs,
                         const int32_t* row_seg_start, int n_rows, int s, int th, int tw, int channels,
                         const float* median, const float* b, int H, int W, float eps, float lo, float hi, int c,
                         int d, int64_t* pad_scratch, int64_t* codes, int64_t* positions, int64_t* channels_out,
                         int64_t* image_ids, uint8_t* key_pad_mask, void* stream);
/* Decode from codes straight into the inverse GEMM's operand planes: lfq.py:105-134 unpack +
 * PN:167-177 de-normalise + FE:635-653 un-patchify + fp16 hi/lo split (scale 2^4, DC to dc[]).
 * Bit-identical to dcta_lfq_indices_to_codes -> dcta_patchnorm_apply(inverse) -> dcta_unpatchify_split. */
int dcta_decode_codes_split(const int64_t* codes, const int32_t* slot_map, const int32_t* img_sel,
                            int64_t n_img, int channels_n, int th, int tw, int p, int rows, int cols,
                            int64_t ld, int out_h, int out_w, const float* median, const float* b, int H,
                            int W, float eps, int c, int d, float scale, void* y_hi, void* y_lo,
                            float* dc, void* stream);

/* ------------------------------------------------------------------ LFQ training terms at scale */
/* UT:355-387 compute_entropy_loss applied to LFQ's distance (LFQ:191) WITHOUT the (n_tok, c, 2^d) tensor: with a +-s
 * codebook the softmax over the 2^d sign patterns is a product of d Bernoulli distributions, p_i(1) = sigmoid(-4 s x_i / T),
 * so the per-sample entropy is a sum of d binary entropies and the average distribution is accumulated as a sum of
 * outer products of the 2^ceil(d/2) and 2^floor(d/2) half-products (csrc/lfq_entropy.cu).  d <= 14.
 *   x (n_tok, c*d) fp32: the quantiser's (projected) input; mask (n_tok) uint8, 1 = valid token;
 *   partial_scratch: dcta_lfq_entropy_ctas() * (2^d + 2) floats;
 *   tables (2 * 2^d): [avg_probs | d avg_entropy / d avg_probs], kept for the backward pass;
 *   result (4): [loss, valid tokens, sample entropy, avg entropy].
 * dcta_lfq_entropy_factorized_backward: grad_x (n_tok, c*d) = grad_out[0] * d loss / d x (zero on masked tokens).
 * dcta_lfq_commit_backward: gradient of dcta_lfq_commit_loss (LFQ:195-200): grad_out[0] * 2 (x - q) / (n_valid c d). */
int dcta_lfq_entropy_ctas(void);
/* d = 14: the forward contraction runs on tcgen05 (fp16 hi / lo operands generated in shared memory, fp32 accumulation in
 * TMEM flushed every 1024 pairs); dcta_lfq_entropy_use_tensor_cores(0) selects the fp32 FMA kernel instead (default 1). */
int dcta_lfq_entropy_use_tensor_cores(int on);
int dcta_lfq_entropy_factorized(const float* x, const uint8_t* mask, int64_t n_tok, int c, int d, float codebook_scale,
                                float temperature, float eps, float* partial_scratch, float* tables, float* result,
                                void* stream);
int dcta_lfq_entropy_factorized_backward(const float* x, const uint8_t* mask, int64_t n_tok, int c, int d,
                                         float codebook_scale, float temperature, const float* tables, const float* result,
                                         const float* grad_out, float* grad_x, void* stream);
int dcta_lfq_commit_backward(const float* x, const uint8_t* mask, const float* grad_out, float* n_valid_scratch,
                             float* grad_x, int64_t n_tok, int cd, float scale, void* stream);

/* ------------------------------------------------------------------ model glue */
/* The row-wise pieces of modeling_dct_autoencoder.py:60-63 (to_patch_embedding = Linear + LayerNorm(eps 1e-4)),
 * :101-112 ((channel, h, w) position embeddings), :85-88 (proj_out = LayerNorm + Linear) on packed token rows; the
 * contractions themselves go through dcta_gemm_split.
 *
 * dcta_ln_pos_rows: out[t] = LN(x[t]) * gamma + beta   [gamma/beta nullable together: no LayerNorm]
 *                            + bias                     [nullable]
 *                            + pos_h[positions[t,0]] + pos_w[positions[t,1]] + pos_c[channels[t]]   [nullable together]
 *   x, out (n_rows, f) fp32, out must not alias x; biased variance, eps inside the square root (torch layer_norm).
 * dcta_split_rows_rowscale: rows (n_rows, d) [-> LayerNorm when gamma given] -> fp16 hi/lo planes (n_rows, ld) of
 *   y * s_row with a power-of-two scale per row (max|y| * s_row in [2^9, 2^10)), and row_scale[t] = post / s_row:
 *   the `row_scale` argument of dcta_gemm_split (post = 1 / the scale of the other operand).  lo may be NULL
 *   (single-precision operand: the hi plane alone). */
int dcta_ln_pos_rows(const float* x, const float* gamma, const float* beta, float eps, const float* bias,
                     const float* pos_c, const float* pos_h, const float* pos_w, const int64_t* channels,
                     const int64_t* positions, float* out, int64_t n_rows, int f, void* stream);
int dcta_split_rows_rowscale(const float* x, const float* gamma, const float* beta, float eps, void* hi, void* lo,
                             float* row_scale, float post, int64_t n_rows, int d, int64_t ld, void* stream);
/* The same operand split with PatchNorm.forward (frozen statistics, patchnorm.py:157-165) applied to the row first:
 *   x (n_rows, d = z) un-normalised patches, channels (n_rows) / positions (n_rows, 2) int64 of the packed batch
 *   (padding rows read the statistics at (0, 0, 0), like the reference), median / b (C, H, W, z), clamp [clamp_lo, clamp_hi].
 *   row_src [nullable] (n_rows) int32: operand row t is row row_src[t] of x, or a row of zeros where it is -1 -- x is then
 *   the TOKEN GRID and row_src the output of dcta_pack_tiles_index, i.e. the gather of FE:516-605 happens in this pass.
 *   The planes equal dcta_split_rows_rowscale(dcta_patchnorm_apply(packed patches)) bit for bit; neither the packed nor
 *   the normalised patches are written.  Needs d % 4 == 0, d <= 1024 and 16-byte aligned rows / tables. */
int dcta_split_rows_patchnorm(const float* x, const int32_t* row_src, const int64_t* channels, const int64_t* positions,
                              const float* median, const float* b, int C, int H, int W, float eps, float clamp_lo,
                              float clamp_hi, void* hi, void* lo, float* row_scale, float post, int64_t n_rows, int d,
                              int64_t ld, void* stream);

/* ------------------------------------------------------------------ VectorQuantize -------- */
/* VQ:29-33 cdist + VQ:467-469 argmax(-dist) + VQ:222-226/477 gather, never materialising the
 * (n_tok, n_codes) distance matrix.
 *   x (n_tok, d); embed (n_codes, d); e2 (n_codes) scratch (sum of squares, filled by the call);
 *   indices (n_tok) i64; quantized (n_tok, d) [nullable] = embed[indices]. */
int dcta_vq_nearest(const float* x, const float* embed, float* e2, int64_t* indices,
                    float* quantized, int64_t n_tok, int n_codes, int d, void* stream);

/* out[i] = sum_k x[i,k]^2 for n rows of d floats. */
int dcta_row_sumsq(const float* x, float* out, int64_t n, int d, void* stream);
/* Same contract as dcta_vq_nearest on tensor cores, in two passes (csrc/vq_tc.cu): approximate distances
 * |e_n|^2 + row_alpha[t] * (x_hi[t] . e_hi[n]) from ONE fp16 tcgen05.mma per product, reduced in the epilogue to the two
 * smallest per token and half of the 128-code slices (cand (n_tok, 4) int32, 16-byte aligned), then an exact fp32
 * re-rank of those candidates with the reference's formula and first-index rule.
 *   x_hi (n_tok, ld): fp16 of x scaled per row by a power of two (dcta_split_rows_rowscale, lo = NULL);
 *   row_alpha (n_tok) = -2 / (row scale * codebook scale) (its row_scale output with post = -2 / codebook scale);
 *   e_hi (n_codes, ld): fp16 of embed * codebook scale; e2 = dcta_row_sumsq(embed) (16-byte aligned).
 * DCTA_ERR_UNSUPPORTED when a d-wide token operand does not fit in shared memory (d > ~700): use dcta_vq_nearest. */
int dcta_vq_nearest_tc(const float* x, const void* x_hi, const float* row_alpha, const float* embed,
                       const void* e_hi, const float* e2, int32_t* cand, int64_t* indices, float* quantized,
                       int64_t n_tok, int n_codes, int d, int64_t ld, void* stream);
/* The same with VQ:1043-1048 folded into the gather: keep [nullable] (n_tok) bytes, 0 = padding token whose row of
 * `quantized` is its own input row x[t] (torch.where(mask, quantize, orig_input) of a projection-free layer); indices are
 * computed for every token either way.
 *   cand_val (n_tok, 4) fp32 scratch + e2_max (one device float = max_n e2[n]) [nullable together]: the first pass also
 *   stores the candidates' approximate values, and the second pass skips the exact re-scoring of a token whose best
 *   candidate leads the runner-up by more than twice the rigorous error bound 2^-8 |x| max|e| (it is then the exact
 *   argmin): same indices, a quarter of the gather traffic for such tokens. */
int dcta_vq_nearest_tc_masked(const float* x, const void* x_hi, const float* row_alpha, const float* embed,
                              const void* e_hi, const float* e2, int32_t* cand, float* cand_val, const float* e2_max,
                              const uint8_t* keep, int64_t* indices, float* quantized, int64_t n_tok, int n_codes, int d,
                              int64_t ld, void* stream);
/* Antialiased bilinear resize of n_planes planes (ih, iw) -> (oh, ow) fp32: the `crop` step of the reference's loader
 * (dataset.py:59-73, torchvision Resize(antialias=True) on a float tensor = F.interpolate(mode="bilinear",
 * antialias=True, align_corners=False)).  _u8: 8-bit input pixels read as u / 255. */
int dcta_resize_bilinear_aa(const float* in, float* out, int64_t n_planes, int ih, int iw, int oh, int ow, void* stream);
int dcta_resize_bilinear_aa_u8(const uint8_t* in, float* out, int64_t n_planes, int ih, int iw, int oh, int ow, void* stream);

/* VectorQuantize codebook learning (VQ:180-220 k-means, VQ:479-500 EMA update).
 *   dcta_vq_cluster_stats: counts (n_codes) = tokens per code, sums (n_codes, d) = sum of their vectors, over the
 *     tokens with mask[t] != 0 (mask uint8, nullable); both outputs are zeroed first (reduce-ready: all-reduce them
 *     across ranks before the update, as the reference does).
 *   dcta_vq_ema_update: cluster_size <- lerp(cluster_size, counts, 1-decay), embed_avg <- lerp(embed_avg, sums, 1-decay),
 *     embed <- embed_avg / (laplace_smoothing(cluster_size, n_codes, eps) * sum(cluster_size)); total_scratch: 1 float.
 *   dcta_vq_kmeans_means: means[c] <- sums[c] / counts[c] for non-empty clusters. */
int dcta_vq_cluster_stats(const float* x, const int64_t* indices, const uint8_t* mask, int64_t n_tok, int d,
                          int n_codes, float* counts, float* sums, void* stream);
int dcta_vq_ema_update(float* embed, float* cluster_size, float* embed_avg, const float* counts, const float* sums,
                       int n_codes, int d, float decay, float eps, float* total_scratch, void* stream);
int dcta_vq_kmeans_means(float* means, const float* counts, const float* sums, int n_codes, int d, void* stream);

/* VectorQuantize commitment loss with the straight-through estimator (VQ:944-952, 976-1003), forward and backward in one
 * flat pass each.
 *   dcta_masked_mse: result[0] = mean over the tokens with mask[t] != 0 (mask uint8 (n_tok), nullable = all) and their dim
 *     elements of (q - x)^2, result[1] = number of those tokens; scratch: dcta_masked_mse_scratch_floats() floats.
 *   dcta_masked_mse_backward: grad_x = grad_q (nullable = 0) + grad_loss[0] * 2 (x - q) * mask / (result[1] * dim). */
int dcta_masked_mse_scratch_floats(void);
int dcta_masked_mse(const float* x, const float* q, const uint8_t* mask, int64_t n_tok, int dim, float* scratch,
                    float* result, void* stream);
int dcta_masked_mse_backward(const float* x, const float* q, const uint8_t* mask, const float* result,
                             const float* grad_loss, const float* grad_q, float* grad_x, int64_t n_tok, int dim,
                             void* stream);

/* ------------------------------------------------------------------ un-patchify ----------- */
/* FE:619-643: slot_map (n_img, channels, th, tw) i32 = flat token index (row*s + slot) of the LAST
 * valid token at that position, -1 where none (the call fills it).
 *   row_img_base (n_rows) i32: global index of the first image of each row. */
int dcta_build_slot_map(const int64_t* channels, const int64_t* positions, const int64_t* image_ids,
                        const uint8_t* key_pad_mask, const int32_t* row_img_base, int n_rows, int s,
                        int64_t n_img, int channels_n, int th, int tw, int32_t* slot_map,
                        void* stream);
/* FE:635-653: planes (n_sel, channels, rows, cols) <- tokens, zero where no token.
 *   img_sel (n_sel) i32 [nullable = identity]: which images of slot_map to render. */
int dcta_unpatchify(const float* patches, const int32_t* slot_map, const int32_t* img_sel,
                    int64_t n_sel, int channels_n, int th, int tw, int p, int rows, int cols,
                    float* planes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCTA_H_ */
