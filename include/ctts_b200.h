/*
 * ctts_b200.h -- C ABI of libctts_b200.so: hand-written sm_100a kernels for the CompTransTTS
 * acoustic-model forward path (reference: keonlee9420/Comprehensive-Transformer-TTS).
 *
 * The reference has no FFI layer: its numerics are PyTorch ATen calls made from Python
 * (SURVEY.md section 8b).  Each entry point below replaces one group of those calls; the
 * reference call site it replaces is cited as file:line (paths relative to the reference root).
 * The Python binding a maintainer would add is a ctypes stub -- see INTEGRATION.md.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`; the caller owns all
 *    memory (PyTorch's allocator in practice); nothing is allocated or retained by the library
 *    except cached TMA descriptors;
 *  - activations are row-major fp32 "token-major" tensors [B, T, C]; lengths are int64 [B]
 *    (the reference's `src_lens` / `mel_lens`); a row t of utterance b is padding iff
 *    t >= lens[b] (utils/tools.py:188-196);
 *  - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no host sync
 *    (the one exception is documented at ctts_length_scan);
 *  - every function returns 0 on success, non-zero on error; ctts_last_error() then returns a
 *    thread-local message.  There is no CPU fallback anywhere.
 */
#ifndef CTTS_B200_H
#define CTTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTTS_ABI_VERSION 1

/* activation selector for ctts_conv1d_gemm */
enum { CTTS_ACT_NONE = 0, CTTS_ACT_RELU = 1, CTTS_ACT_GELU = 2, CTTS_ACT_TANH = 3, CTTS_ACT_SWISH = 4 };

/* arithmetic selector for ctts_conv1d_gemm */
enum {
    CTTS_MATH_FP32 = 0,   /* CUDA-core FP32 FMA: used upstream of every quantiser      */
    CTTS_MATH_BF16X3 = 1  /* tcgen05 kind::f16, operands split bf16 hi+lo, 3 MMAs, FP32 accumulate   */
};

int ctts_abi_version(void);
const char* ctts_last_error(void);
/* compute capability major*10+minor of the current device; the library refuses to run below 100 */
int ctts_device_arch(void);

/* ---- phoneme embedding + sinusoidal positions -------------------------------------------
 * transformer_fs2.py:113-119 (TextEncoder.forward_embedding), blocks.py:85-104, utils/tools.py:640-652.
 *   pos[b,s]  = (# of non-zero tokens in tokens[b,0..s]) if tokens[b,s] != 0 else 0
 *   word      = embed_scale * table[tokens]
 *   x         = (word + pe[pos]) * keep          (keep: s < lens[b]; FFTBlocks.forward :60; lens may be NULL)
 * pos_mode 0: as above (fs2).  pos_mode 1: pos[b,s] = s (absolute; transformer.py:72-74, fastformer.py:62-64,
 * conformer.py:82-84 with embed_scale 1 and the interleaved sin/cos table).
 * pe is the [pe_rows, C] sinusoid table (row 0 = zeros); pe_rows must be > S.  `word` is not masked
 * (it is the reference's second return value, consumed by the aligner).
 */
int ctts_embed_tokens(const int64_t* tokens, const float* table, const float* pe, int pe_rows, float embed_scale,
                      int B, int S, int C, int vocab, float* x, float* word, const int64_t* lens, int pos_mode,
                      void* stream);

/* ---- y = (x + alpha * pe[pos(x[..., 0] != 0)]) [* keep]   (out of place: y != x) -----------------------
 * FFTBlocks.forward transformer_fs2.py:54-60 (decoder positions) and PitchPredictor.forward
 * modules.py:1349-1350.  `alpha` is a device scalar (the learnable pos_embed_alpha).  If lens != NULL
 * rows t >= lens[b] are zeroed afterwards (the `* nonpadding_mask_TB` of :60).
 * pos_mode 1: pos = t (absolute table rows, transformer.py:137-141); alpha may then be NULL (= 1).
 */
int ctts_add_positions(const float* x, const float* pe, int pe_rows, const float* alpha, const int64_t* lens, int B, int T,
                       int C, int pos_mode, float* y, void* stream);

/* ---- y = LayerNorm_C(x) * gamma + beta [* keep] -------------------------------------------
 * blocks.py:137-156 (eps 1e-12), transformer_fs2.py:41,65-66 (final nn.LayerNorm eps 1e-5).
 * x, y: [B*T, C].  lens may be NULL.
 */
int ctts_layernorm(const float* x, const float* gamma, const float* beta, float eps, const int64_t* lens, int B, int T,
                   int C, float* y, void* stream);

/* ---- the dense workhorse: Conv1d(k taps, "same" zero padding) / Linear as an implicit GEMM --------
 *   acc[b,t,n] = sum_{j<taps} sum_{c<Cin} x[b, t + j - taps/2, c] * w[n, j*Cin + c]      (x = 0 outside [0,T))
 *   v          = (acc + bias[n]) * alpha
 *   v          = v * col_scale[n] + col_shift[n]          (folded eval-mode BatchNorm1d)
 *   v          = act(v) + residual[b,t,n]
 *   y[b,t,n]   = lens ? (t < lens[b] ? v : 0) : v
 * bias / col_scale / col_shift / residual / lens may be NULL.  `w` is the PACKED weight
 * [N, taps*Cin] produced by ctts_pack_conv_weight.  Replaces nn.Conv1d / nn.Linear at:
 * transformer_fs2.py:226-238 (FFN), F.multi_head_attention_forward in/out projections (:385-394),
 * modules.py:1299-1310,1343-1356 (predictor stacks), modules.py:140-148 (PostNet), CompTransTTS.py:133
 * (mel_linear), modules.py:1195-1196 (aligner projections).
 * math = CTTS_MATH_FP32: Cin % 16 == 0.  math = CTTS_MATH_BF16X3: see ctts_gemm_bf16x3.
 */
int ctts_conv1d_gemm(const float* x, const float* w, const float* bias, float alpha, const float* col_scale,
                     const float* col_shift, int act, const float* residual, const int64_t* lens, int B, int T,
                     int Cin, int N, int taps, float* y, void* stream);

/* torch Conv1d weight [N, Cin, taps] -> packed [N, taps*Cin] (tap-major).  taps == 1: plain copy. */
int ctts_pack_conv_weight(const float* w, int N, int Cin, int taps, float* packed, void* stream);

/* ---- masked multi-head self-attention (softmax(q k^T * scale + key_padding_mask) v) ---------------
 * transformer_fs2.py:385-394 -> F.multi_head_attention_forward; transformer.py:233-252.
 * qkv: [B, T, 3*C] (q | k | v, heads contiguous inside each C), head_dim = C / H in {32, 64, 128}.
 * Keys s >= lens[b] are excluded; query rows t >= lens[b] produce zeros (they are zeroed by the block's
 * pad mask in the reference, transformer_fs2.py:192).  Scores are never materialised.
 */
int ctts_attention(const float* qkv, const int64_t* lens, int B, int T, int C, int H, float scale, float* out,
                   void* stream);

/* ---- duration decoding: clamp(round(exp(logd) - 1) * d_control, min 0)   modules.py:1060-1063 ---- */
int ctts_decode_durations(const float* log_d, float d_control, int n, float* dur, void* stream);

/* ---- LengthRegulator scan (integer; bit-exact) -----------------------------------------
 * modules.py:1222-1249 + utils/tools.py:598-628.  Either dur_f32 or dur_i64 is non-NULL, [B, S].
 *   reps[b,j]   = max(int(dur[b,j]), 0)                 (LR: truncation toward zero)
 *   cum_lr[b,j] = sum_{i<=j} reps[b,i]        mel_len[b] = cum_lr[b,S-1]
 *   cum_m2p[b,j]= sum_{i<=j} round(dur[b,i]) * (j < src_lens[b])    (dur_to_mel2ph: half-to-even rounding)
 * cum_lr / cum_m2p: int32 [B, S]; mel_len: int64 [2*B]: [0,B) = LR lengths (the reference's mel_len),
 * [B,2B) = sum of the rounded durations (the length of mel2ph).  The caller reads the maxima back to size the
 * output when the reference's max_len is None -- the single host sync of the inference path
 * (the reference has B*S of them, modules.py:1241).
 */
int ctts_length_scan(const float* dur_f32, const int64_t* dur_i64, const int64_t* src_lens, int B, int S,
                     int32_t* cum_lr, int32_t* cum_m2p, int64_t* mel_len, void* stream);

/* ---- LengthRegulator expand ----------------------------------------------------------
 *   j = the phoneme with cum_lr[b,j-1] <= t < cum_lr[b,j]
 *   row = table ? table[row_index[b,j]] : src[b,j]          ([*, C] fp32)
 *   out[b,t,:] = (accumulate ? out[b,t,:] : 0) + (j exists ? row : 0)
 * mel2ph (nullable, int64 [B, M2]) = 1-based index from cum_m2p, 0 beyond the end (utils/tools.py:627).
 * Used for x (modules.py:1064) and, with table = energy_embedding.weight and row_index = bucket ids,
 * for the phoneme-level energy embedding (modules.py:1099).
 */
int ctts_length_expand(const float* src, const float* table, const int64_t* row_index, const int32_t* cum_lr,
                       int B, int S, int C, int M, int accumulate, float* out, const int32_t* cum_m2p,
                       int64_t* mel2ph, int M2, void* stream);

/* ---- pitch: CWT spectrogram -> f0 -> mel-scale bucket -> (optional) embedding add ---------------
 * modules.py:907-938 + utils/pitch_tools.py:258-294 (cwt2f0_norm), :69-82 (denorm_f0), :27-36 (f0_to_coarse).
 *   rec[b,t] = sum_i cwt[b,t,i] * scale_w[i]   (i < 10; cwt row stride = cwt_stride floats)
 *   rec      = (rec - mean_t rec) / std_t rec  (over all T frames, unbiased)
 *   f0n      = log2(exp(rec * (std[b]*std_scale) + mean[b]) + eps)
 *   uv       = uv_src ? (uv_src[b,t] > 0) : (cwt[b,t,10] > 0)   (use_uv; uv_from_cwt selects)
 *   f0d      = uv ? 0 : 2^f0n                -> f0_denorm[b,t];  f0_norm (nullable) receives f0n
 *   idx      = f0_to_coarse(f0d)             -> pitch_idx[b,t] (int64)
 * stats: [B,2] rows (mean, std) with element stride given by stats_stride (2 for the MLP output,
 * or separate arrays via mean/std pointers).
 */
int ctts_cwt_to_pitch(const float* cwt, int cwt_stride, const float* scale_w, const float* mean, const float* std,
                      int stat_stride, float std_scale, float eps, const float* uv_src, int use_uv, int B, int T,
                      float* f0_norm, float* f0_denorm, int64_t* pitch_idx, void* stream);

/* f0 given (teacher forcing): f0_denorm = uv ? 0 : 2^f0 ; idx = coarse(f0_denorm).  modules.py:933-938 */
int ctts_f0_to_pitch(const float* f0_norm, const float* uv_src, int n, float* f0_denorm, int64_t* pitch_idx,
                     void* stream);

/* pitch_type 'frame' / 'ph' (modules.py:890-906,927-938; preprocess.yaml pitch_type):
 * ctts_frame_pitch    per frame: f0 = f0_target ? f0_target : pred[i*ldp]; uv = uv_target > 0 or pred[i*ldp+1] > 0 (use_uv);
 *                     padding = mel2ph == 0; f0_denorm = (uv | padding) ? 0 : 2^f0; f0_out = padding ? 0 : f0 (may alias the
 *                     target: the reference zeroes it in place; without a target pred[i*ldp] itself is zeroed at padded
 *                     frames, because f0 is a view of pitch_pred there); idx = f0_to_coarse(f0_denorm)
 * ctts_gather_index   out[b,t] = mel2ph[b,t] > 0 ? idx_ph[b, mel2ph[b,t]-1] : 0      (F.pad + torch.gather, :900-901)
 * ctts_phoneme_pitch  get_phoneme_level_pitch (modules.py:874-880, utils/tools.py:47-53): per-phoneme mean of frame f0 */
int ctts_frame_pitch(float* pred, int ldp, const float* f0_target, const float* uv_target, const int64_t* mel2ph, int use_uv,
                     int n, float* f0_out, float* f0_denorm, int64_t* pitch_idx, void* stream);
int ctts_gather_index(const int64_t* idx_ph, const int64_t* mel2ph, int B, int S, int M, int64_t* out, void* stream);
int ctts_phoneme_pitch(const float* f0, const int64_t* mel2ph, const int64_t* src_lens, const int64_t* mel_lens, int B, int S,
                       int M, float* out, void* stream);

/* x[r,:] += table[idx[r],:]   (nn.Embedding + add; modules.py:938,1089) */
int ctts_gather_add(const float* table, const int64_t* idx, int rows, int C, int table_rows, float* x, void* stream);

/* idx[i] = #{ bins[k] < v[i] }  == torch.bucketize(v, bins) (right=False); modules.py:954-958 */
int ctts_bucketize(const float* v, float v_scale, const float* bins, int n_bins, int n, int64_t* idx, void* stream);

/* y = x + spk[b] broadcast over T (modules.py:985-988) */
int ctts_add_row_broadcast(const float* x, const float* row, int B, int T, int C, float* y, void* stream);

/* ---- liu2021 implicit prosody predictors (eval path; modules.py:572-648, 1002-1023) ----------------------
 * ctts_gru_bidir      recurrence of a bidirectional single-layer nn.GRU (gates r|z|n).  gi_* = x W_ih^T + b_ih for every
 *                     step ([B,T,3H], produced by ctts_conv1d_gemm); out [B,T,2H] (fwd | bwd), h_final [B,2H].  Runs over
 *                     all T (padded) steps like the reference (no packing).  3H <= 1024; W_hh^T lives in shared memory.
 * ctts_linear_smallk  y = x W^T + b (+ residual) for tiny K (the 4-d phoneme prosody code -> 256, modules.py:861).
 */
int ctts_gru_bidir(const float* gi_fwd, const float* gi_bwd, const float* w_hh_fwd, const float* b_hh_fwd,
                   const float* w_hh_bwd, const float* b_hh_bwd, int B, int T, int H, float* out, float* h_final,
                   void* stream);
int ctts_linear_smallk(const float* x, const float* w, const float* bias, const float* residual, int rows, int K, int N,
                       float* y, void* stream);

/* ---- unsupervised duration modelling (learn_alignment: True) --------------------------------------
 * ctts_aligner_attention  AlignmentEncoder.forward score assembly, modules.py:1198-1212.  q [B,M,C] / k [B,S,C] are the
 *   projected mel / text features (the conv stacks run through ctts_conv1d_gemm), prior [B,S,M] is the caller's
 *   attn_priors (read transposed).  soft, logprob: [B,M,S] (= the reference's [B,1,M,S]).  The first log_softmax runs
 *   over ALL S columns, the mask is applied only before the second softmax (reference quirk, kept).
 * ctts_mas                 binarize_attention_parallel / b_mas / mas_width1, modules.py:36-75,863-872: monotonic Viterbi
 *   path per utterance on log(attn); hard [B,M,S] 0/1, dur [B,S] = hard.sum(M).  prev_workspace: B*M*S bytes.
 *   Replaces the reference's D2H copy + numba CPU loop + H2D copy per training step.
 * ctts_phoneme_energy      get_phoneme_level_energy, modules.py:882-888 + utils/tools.py:56-66 (sequential, in place).
 *   workspace: B*M floats; out [B,S].
 */
int ctts_aligner_attention(const float* q, const float* k, const float* prior, const int64_t* src_lens, float temperature, int B,
                           int M, int S, int C, float* soft, float* logprob, void* stream);
int ctts_mas(const float* attn, const int64_t* src_lens, const int64_t* mel_lens, int B, int M, int S, uint8_t* prev_workspace,
             float* hard, float* dur, void* stream);
int ctts_phoneme_energy(const float* dur, const int64_t* src_lens, const float* energy, int B, int S, int M, float* workspace,
                        float* out, void* stream);

/* ---- FP32 batched GEMM with explicit strides (conformer attention products) ------------
 *   y[zo,zh][t, n] = alpha * sum_k x[zo,zh][t, k] * w[zo,zh][n, k]        z = zo*mod + zh in [0, Z)
 * row pointers: x + zo*x_so + zh*x_sh + t*x_ld, w + zo*w_so + zh*w_sh + n*w_ld, y + zo*y_so + zh*y_sh + t*y_ld.
 * K % 16 == 0.  lens (nullable): rows t >= lens[z / lens_div] are written as zeros.
 */
int ctts_batched_gemm_fp32(const float* x, const float* w, float alpha, const int64_t* lens, int lens_div, int Z, int mod,
                           int T, int K, int N, long long x_so, long long x_sh, int x_ld, long long w_so, long long w_sh,
                           int w_ld, long long y_so, long long y_sh, int y_ld, float* y, void* stream);

/* ---- fastformer additive attention pooling (fastformer.py:308-322, 326-336) ------------------------
 *   s[t,h] = logits[b,t,h] / sqrt(head_size) + (t < lens[b] ? -10000 : 0)   -- the reference's inverted mask is kept
 *   pooled[b, h*head_size + e] = sum_t softmax_t(s)[t,h] * values[b, t, h*head_size + e]
 * logits [B,T,heads], values [B,T,heads*head_size], pooled [B, heads*head_size]; head_size <= 4.
 */
int ctts_fastformer_pool(const float* logits, const float* values, const int64_t* lens, int B, int T, int heads, int head_size,
                         float* pooled, void* stream);

/* y = a (+|*) b, optionally pad-masked.  op 0 add, 1 multiply; b_rowwise: b is [B, C] broadcast over T. */
int ctts_binary(const float* a, const float* b, int op, int b_rowwise, const int64_t* lens, int B, int T, int C, float* y,
                void* stream);

/* ---- conformer convolution module, elementwise stages (conformer.py:459-469) -----------------------
 * ctts_glu:            g[r, c] = h[r, c] * sigmoid(h[r, C + c])                    h: [rows, 2C]
 * ctts_dwconv_bn_swish y[b,t,c] = swish(scale[c] * sum_j g[b,t+j-K/2,c] w[c,j] + shift[c])   (eval BatchNorm1d folded)
 */
int ctts_glu(const float* h, int rows, int C, float* g, void* stream);
int ctts_dwconv_bn_swish(const float* g, const float* w, int K, const float* scale, const float* shift, int B, int T, int C,
                         float* y, void* stream);

/* ---- conformer relative-position attention, score assembly (conformer.py:405-431) ------------------
 *   P[z,i,:] = softmax_j( (content[z,i,j] + shift(pos)[z,i,j]) / sqrt_dim )   -- no padding mask (quirk kept)
 *   shift(pos)[i,j] = j <= i ? pos[i, T-1-i+j] : (j == i+1 ? 0 : pos[i+1, j-i-2])
 * content, pos: [Z, T, T]; P: [Z, T, ldp] (zero padded up to ldp).
 */
int ctts_relshift_softmax(const float* content, const float* pos, int Z, int T, int ldp, float sqrt_dim, float* P,
                          void* stream);

/* tensor-core form of the conformer attention (decoder): the same score assembly reading content / pos with row stride ld
 * and writing P as bf16 planes [Z, T, ldp] (the A operand of the P.V GEMM; fp32 probabilities never reach HBM), and the
 * head re-layout that lets the 32-wide heads of conformer.py:375-380 ride the 64-wide k-blocks of the GEMM engine:
 *   planes[r, h*DHp + d] = d < DH ? x[r, c0 + h*DH + d] + bias[h*DH + d] : 0     (bias = u_bias / v_bias or NULL) */
int ctts_relshift_softmax_planes(const float* content, const float* pos, int Z, int T, int ld, int ldp, float sqrt_dim,
                                 int n_planes, void* const* planes, void* stream);
int ctts_pad_heads_planes(const float* x, const float* bias, int rows, int ld_in, int c0, int H, int DH, int DHp, int n_planes,
                          void* const* planes, void* stream);

/* x[b, t, c0 + h*DH + d] -> xt[(b*H + h), d, t]  (row stride ldt >= T, zero padded): K-major operand for P.V */
int ctts_transpose_heads(const float* x, int B, int T, int ld_in, int c0, int H, int DH, int ldt, float* xt, void* stream);

/* ---- tcgen05 tensor-core GEMM (the decoder / PostNet engine) ---------------------------------
 * Same contract as ctts_conv1d_gemm but the operands are bf16 hi/lo planes:
 *   x_hi, x_lo : [B, T, Cin] bf16, x ~= x_hi + x_lo     (written by the producing kernel's epilogue)
 *   w_hi, w_lo : [N, taps*Cin] bf16 packed planes (ctts_split_bf16 of the packed fp32 weight)
 *   acc = x_hi*w_hi + x_hi*w_lo + x_lo*w_hi   (three tcgen05.mma per k-slice, FP32 accumulate in TMEM)
 * The epilogue is identical; additionally y_hi / y_lo (nullable) receive the bf16 split of y so
 * the next GEMM can consume it without a separate pass.  Requirements: Cin % 64 == 0, N % 16 == 0.
 * Tiles are fetched with TMA from 3-D tensor maps over [B, T, Cin]; the conv halo (rows t < 0 or
 * t >= T) is produced by TMA out-of-bounds zero fill.
 */
int ctts_gemm_bf16x3(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                     float alpha, const float* col_scale, const float* col_shift, int act, const float* residual,
                     const int64_t* lens, int B, int T, int Cin, int N, int taps, float* y, void* y_hi, void* y_lo,
                     void* stream);

/* ---- tensor-core masked self-attention (decoder) ---------------------------------------------
 * Same function as ctts_attention on bf16 hi/lo planes of qkv [B, T, 3C], head_dim % 64 == 0.  Four launches:
 *   S = scale * q k^T      (tcgen05 bf16x3, batched over (b, head); scores [B*H, T, Tp] fp32, Tp = T rounded up to 8)
 *   Vt = V^T planes        ([B*H, head_dim, Tp])
 *   P = softmax over keys < lens[b] of S, as bf16 hi/lo planes (zeros for masked keys / padded query rows)
 *   out = P V              (tcgen05 bf16x3) -> out_hi/out_lo planes [B, T, C] and/or out_f32; rows t >= lens[b] are zero
 * Workspaces are the caller's: scores B*H*T*Tp floats, p_hi/p_lo B*H*T*Tp bf16 each, vt_hi/vt_lo B*C*Tp bf16 each.
 * Replaces transformer_fs2.py:385-394 / transformer.py:233-252 on the decoder (downstream of every quantiser).
 */
int ctts_attention_bf16x3(const void* qkv_hi, const void* qkv_lo, const int64_t* lens, int B, int T, int C, int H,
                          float scale, float* scores, void* p_hi, void* p_lo, void* vt_hi, void* vt_lo, void* out_hi,
                          void* out_lo, float* out_f32, void* stream);

/* ---- generalisation to n_planes bf16 planes per operand ------------------------------------------------
 * n_planes 2: x ~= p0 + p1 (16 mantissa bits), 3 MMAs per k-slice ("bf16x3", decoder / PostNet);
 * n_planes 3: x ~= p0 + p1 + p2 (24 mantissa bits), 6 MMAs per k-slice ("bf16x6": FP32-equivalent; measured 3e-6 on the
 *             whole model, the same as FP32 summation-order noise) -- used for everything UPSTREAM of a quantiser
 *             (encoder, duration / pitch / energy predictors), SURVEY.md H1.
 * x_planes / w_planes / y_planes are HOST arrays of n_planes device pointers.  Otherwise as ctts_gemm_bf16x3.
 */
int ctts_gemm_split(int n_planes, const void* const* x_planes, const void* const* w_planes, const float* bias, float alpha,
                    const float* col_scale, const float* col_shift, int act, const float* residual, const int64_t* lens,
                    int B, int T, int Cin, int N, int taps, float* y, void* const* y_planes, void* stream);
/* GEMM + residual + LayerNorm in one launch (2 operand planes, N = 256): y = (conv(x) + bias) * alpha + residual with rows
 * t >= lens[b] zeroed (y may alias residual); ln_planes (and ln_y, nullable) = LayerNorm(y) * ln_gamma + ln_beta over the 256
 * channels, zeroed for t >= lens[b] when ln_masked != 0.  Replaces a projection followed by nn.LayerNorm in an FFT block:
 * transformer_fs2.py:176-200 (self_attn.out_proj -> layer_norm2, ffn.ffn_2 -> the next block's layer_norm1 or the stack's
 * final layer_norm, :60-66). */
int ctts_gemm_split_ln(const void* const* x_planes, const void* const* w_planes, const float* bias, float alpha,
                       const float* residual, const int64_t* lens, int B, int T, int Cin, int N, int taps, float* y,
                       const float* ln_gamma, const float* ln_beta, float ln_eps, int ln_masked, float* ln_y,
                       void* const* ln_planes, void* stream);
int ctts_split_planes(const float* x, size_t n, int n_planes, void* const* planes, void* stream);
int ctts_layernorm_planes(const float* x, const float* gamma, const float* beta, float eps, const int64_t* lens, int B, int T,
                          int C, float* y, int n_planes, void* const* planes, void* stream);

/* n-plane generalisation of ctts_attention_bf16x3 (n_planes 3 = FP32-equivalent, used by the encoder).  Plane arguments are
 * HOST arrays of n_planes device pointers; workspaces as above, per plane. */
int ctts_attention_split(int n_planes, const void* const* qkv_planes, const int64_t* lens, int B, int T, int C, int H,
                         float scale, float* scores, void* const* p_planes, void* const* vt_planes, void* const* out_planes,
                         float* out_f32, void* stream);

/* Fused self-attention for short sequences (T <= 128, head_dim 128, 3 planes = FP32-equivalent): one CTA per (batch, head)
 * keeps Q, K, the scores and the probability planes on the SM and reads V straight from the qkv planes (MN-major operand)
 * -- ONE launch instead of the four of ctts_attention_split.  Output planes [B, T, C]; rows t >= lens[b] are zero.
 * Replaces transformer_fs2.py:385-394 (F.multi_head_attention_forward) on the encoder at LJSpeech phoneme lengths. */
int ctts_attention_small(const void* const* qkv_planes, const int64_t* lens, int B, int T, int C, int H, float scale,
                         void* const* out_planes, void* stream);

/* ---- fused ("flash") tensor-core self-attention, head_dim 128, 2 planes ------------------------------------
 * One CTA per (batch*head, 128 queries): S = Q K^T in TMEM, softmax in registers, probability planes in shared memory,
 * O = P V in TMEM; the keys are swept twice (row statistics, then P V) so nothing is rescaled and no score ever reaches
 * HBM.  V is read from the qkv planes ([keys][dims] = the MN-major form of the B operand); vt_hi / vt_lo are vestigial
 * (ignored, may be NULL).  Output planes [B, T, C]; rows t >= lens[b]
 * are zero.  Same function as ctts_attention_bf16x3 (transformer_fs2.py:385-394, transformer.py:233-252).
 */
int ctts_flash_attention_bf16x3(const void* qkv_hi, const void* qkv_lo, const void* vt_hi, const void* vt_lo,
                                const int64_t* lens, int B, int T, int C, int H, float scale, void* out_hi, void* out_lo,
                                void* stream);
/* qkv planes [B, T, 3C] -> V^T planes [B*H, C/H, Tp] (Tp = T rounded up to 8, zero padded) */
int ctts_transpose_v_planes(int n_planes, const void* const* qkv_planes, int B, int T, int C, int H, void* const* vt_planes,
                            void* stream);

/* development aid: when non-NULL, every CTA of the following ctts_gemm_split launches stores four clock64() stamps
 * {start, setup done, accumulator ready, epilogue done} at device_buffer[4 * cta]; used by profiles/ scripts only. */
int ctts_debug_set_timing_buffer(long long* device_buffer);

/* fp32 -> (bf16 hi, bf16 lo) with hi = rn(x), lo = rn(x - hi) */
int ctts_split_bf16(const float* x, size_t n, void* hi, void* lo, void* stream);

/* LayerNorm whose output is written as fp32 (nullable) and as bf16 hi/lo planes */
int ctts_layernorm_split(const float* x, const float* gamma, const float* beta, float eps, const int64_t* lens, int B,
                         int T, int C, float* y, void* y_hi, void* y_lo, void* stream);


/* =====================================================================================================================
 * TRAINING STEP (reference: train.py:104-123 -- forward in model.train(), loss.backward()).
 * The reference's backward is torch.autograd over ATen ops; here every backward step is an explicit entry point.  The
 * host side (ctts_b200/train_engine.py) records the forward on a tape and replays it in reverse.  All gradient outputs
 * that belong to PARAMETERS accumulate (+=) into the caller's flat gradient arena; activation gradients take an explicit
 * `accumulate` flag.  Dense contractions: dgrad = ctts_gemm_split on weights re-packed by ctts_pack_conv_weight_dgrad,
 * wgrad = ctts_gemm_wgrad; ctts_gemm_generic is the FP32 CUDA-core GEMM for the odd shapes (1/2/4/11-wide heads,
 * aligner products) and the implementation the tensor-core wgrad is tested against.
 * ===================================================================================================================== */

/* y[z][m,n] = alpha * sum_k A[z][m,k] * B[z][n,k'] (+ y[z][m,n] if accumulate),  z = zo*zmod + zi.
 * Fully strided: a_str = {zo, zi, m, k, kb}, b_str = {zo, zi, n, k, kb}, y_str = {zo, zi, m, n} (element strides).
 * The reduction index splits as k = kb*Kin + kt (Kin <= 0: Kin = K); B is read at kt + shift0 + z*shift_z and is zero
 * outside [0, Kin): the tap shift of a Conv1d weight gradient (reduction over (utterance, time)). */
int ctts_gemm_generic(const float* a, const float* b, float* y, int Z, int zmod, int M, int N, int K, const long long* a_str,
                      const long long* b_str, const long long* y_str, int Kin, int shift0, int shift_z, float alpha,
                      int accumulate, void* stream);

/* Backward of the GEMM epilogue  v = (acc + bias) * alpha; y = act(v) [* keep]:
 *   dz[r,n] = dy[r,n] * keep(r) * act'(.) * alpha;  dbias[z, n] += sum_r dz[r,n]
 * `ref` = pre-activation v for GELU / SWISH, the OUTPUT y for RELU / TANH, unused for NONE.  dy/ref/dz: [Z, rows, N];
 * keep(r): global row z*rows + r = (b, t) with t < lens[b] (lens nullable).  dz may alias dy; dz or dbias may be NULL.
 * With Z = B, rows = T, dz = NULL it is the backward of a row broadcast (modules.py:985-988): dbias[b, n] += sum_t dy. */
int ctts_act_bwd(const float* dy, const float* ref, int act, float alpha, const int64_t* lens, int Z, int T, int rows, int N,
                 float* dz, float* dbias, void* stream);

/* ctts_act_bwd fused with the operand preparation of the two GEMMs that consume dz: writes dz (fp32, nullable, may alias dy),
 * its row-major bf16 planes [B, T, N] (dgrad operand) and its time-major planes [B, N, Tp] (wgrad operand), adds the column
 * sums to dbias.  N % 4 == 0. */
int ctts_act_bwd_planes(const float* dy, const float* ref, int act, float alpha, const int64_t* lens, int B, int T, int N, int Tp,
                        float* dz, int n_planes, void* const* dz_planes, void* const* dzT_planes, float* dbias, void* stream);
/* y = (res + dropout(x)) * keep: residual add behind a dropout (transformer_fs2.py:190-192,197-199) in one pass; the mask is
 * the one ctts_dropout draws for (seed, offset [+ *offset_dev]) */
int ctts_dropout_add(const float* x, const float* res, const int64_t* lens, int B, int T, int C, float p, unsigned long long seed,
                     unsigned long long offset, const unsigned long long* offset_dev, float* y, void* stream);

/* LayerNorm backward (blocks.py:137-156, nn.LayerNorm): dx (+)= ..., dgamma += , dbeta += ; statistics recomputed from x */
int ctts_layernorm_bwd(const float* x, const float* gamma, const float* dy, float eps, const int64_t* lens, int B, int T,
                       int C, float* dx, int accumulate, float* dgamma, float* dbeta, void* stream);

/* x[b,t,:] = 0 for t >= lens[b]  (backward of the `* nonpadding` masks, transformer_fs2.py:60,192,199) */
int ctts_mask_rows(float* x, const int64_t* lens, int B, int T, int C, void* stream);
/* y = (accumulate ? y : 0) + a * x      (gradient fan-in; predictor_grad scaling modules.py:1026,893) */
int ctts_axpy(const float* x, float a, size_t n, int accumulate, float* y, void* stream);
/* y[r,c] = (accumulate ? y : 0) + a * x[r,c] * s[r] */
int ctts_rowscale_axpy(const float* x, const float* s, float a, int rows, int C, int accumulate, float* y, void* stream);

/* nn.Embedding backward: dtable[idx[r]] += scale * dy[r] (rows t >= lens[b] and idx == skip_idx (padding_idx) skipped) */
int ctts_scatter_add_rows(const float* dy, const int64_t* idx, const int64_t* lens, int T, int rows, int C, int table_rows,
                          int skip_idx, float scale, float* dtable, void* stream);

/* LengthRegulator backward (modules.py:1222-1249): dsrc[b,j] (+)= sum of dy over the frames phoneme j was copied to */
int ctts_length_expand_bwd(const float* dy, const int32_t* cum_lr, int B, int S, int C, int M, int accumulate, float* dsrc,
                           void* stream);

/* gradient of the learnable positional scale: dalpha += sum dy * pe[pos]  (transformer_fs2.py:54-58, modules.py:1349) */
int ctts_add_positions_bwd(const float* dy, const float* x, const float* pe, int pe_rows, const int64_t* lens, int B, int T,
                           int C, int pos_mode, float* dalpha, void* stream);

/* softmax over materialised scores [Z, T, ld] (keys >= lens[z/H] excluded; rows t >= lens zero if mask_rows) and its
 * backward dS = P * (dP - sum P dP) * scale.  FP32 attention backward path (transformer_fs2.py:385-394). */
int ctts_masked_softmax(const float* S, const int64_t* lens, int H, int Z, int T, int Tk, int ld, int mask_rows, float* P,
                        void* stream);
int ctts_softmax_bwd(const float* P, const float* dP, int Z, int T, int Tk, int ld, float scale, float* dS, void* stream);

/* BatchNorm1d in training mode (PostNet modules.py:140-148; conformer.py:465): batch mean / biased variance over all rows
 * (padded frames included, as in the reference), normalise + activation, running-buffer update, backward. */
int ctts_bn_stats(const float* x, int rows, int C, float* mean, float* var, void* stream);
int ctts_bn_act_fwd(const float* x, const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                    int act, int rows, int C, float* y, int n_planes, void* const* planes, void* stream);
int ctts_bn_update_running(const float* mean, const float* var, int rows, float momentum, int C, float* running_mean,
                           float* running_var, int64_t* num_batches_tracked, void* stream);
/* dx = d/dx of act(BN(x)); dgamma += ; dbeta += ; workspace: 2*C floats */
int ctts_bn_bwd(const float* dy, const float* x, const float* mean, const float* var, const float* gamma, const float* beta,
                float eps, int act, int rows, int C, float* dx, float* dgamma, float* dbeta, float* workspace, void* stream);

/* y = act(x) as fp32 and / or bf16 planes: training keeps the pre-activation of GELU / Swish layers for the backward pass
 * (transformer_fs2.py:231-232), so the activation is its own pass there */
int ctts_act_fwd(const float* x, size_t n, int act, float* y, int n_planes, void* const* planes, void* stream);
/* y = sum of n_planes bf16 planes (fp32 value of a tensor the tensor-core kernels emitted as planes only) */
int ctts_merge_planes(int n_planes, const void* const* planes, size_t n, float* y, void* stream);
/* dst[r, 0:C] = (accumulate ? dst : 0) + src[r, 0:C], independent row strides: x_org[:, 0] gather for the CWT statistics
 * MLP (modules.py:909-912) and its scatter in the backward pass; re-striding of attn_soft for the soft upsampling */
int ctts_copy_rows(const float* src, long long src_stride, int rows, int C, float* dst, long long dst_stride, int accumulate,
                   void* stream);

/* Dropout with a counter-based Philox4x32-10 stream: y = x * keep / (1 - p); the mask is a pure function of
 * (seed, offset [+ *offset_dev], element index), so the backward pass calls the same entry on dy.  offset_dev (nullable,
 * device) is a step counter added to `offset`: a training step replayed as a CUDA graph then draws fresh masks.  Replaces
 * F.dropout / nn.Dropout (transformer_fs2.py:58,118,190,197,237; modules.py:144-145,1287,1337). */
int ctts_dropout(const float* x, size_t n, float p, unsigned long long seed, unsigned long long offset,
                 const unsigned long long* offset_dev, float* y, void* stream);

/* weight layouts of the backward GEMMs: wd[c, j*N + n] = w[n, c, taps-1-j] (dgrad operand);
 * dw[n, c, j] (+)= dw_packed[n, j*Cin + c] (wgrad result -> torch Conv1d layout) */
int ctts_pack_conv_weight_dgrad(const float* w, int N, int Cin, int taps, float* wd, void* stream);
int ctts_unpack_conv_wgrad(const float* dw_packed, int N, int Cin, int taps, int accumulate, float* dw, void* stream);

/* x fp32 [Z, R, ld_in] columns [c0, c0+C) -> bf16 planes [Z, taps, C, Rp] (rows contiguous, zero padded) with
 * out[z, tap, c, r] = x[z, r + tap - taps/2, c]: the K-major (and per-tap pre-shifted) operands of ctts_gemm_wgrad */
int ctts_split_transpose(const float* x, int Z, int R, int C, int ld_in, int c0, int Rp, int taps, int n_planes,
                         void* const* planes, void* stream);

/* tcgen05 weight gradient: dw_packed[n, tap*Cin + c] (+)= alpha * sum_{b,t} dz[b,t,n] * x[b, t+tap-taps/2, c] from the
 * transposed planes dzT [B, N, Tp] (ctts_split_transpose, taps 1) and xT [B, taps, Cin, Tp] (ctts_split_transpose, taps) */
int ctts_gemm_wgrad(int n_planes, const void* const* dzT_planes, const void* const* xT_planes, int B, int T, int Tp,
                    int Cin, int N, int taps, float alpha, int accumulate, float* dw_packed, void* stream);

/* The same weight gradient straight from the ROW-MAJOR planes dz [B, T, N] and x [B, T, Cin] (both MN-major operands of
 * the tensor core, the tap is a row offset of the x box): no transposed copies, one set of x planes for all taps.  Needs
 * Cin % 128 == 0 and N % 8 == 0.  Backward of nn.Conv1d / nn.Linear weights (autograd of model/transformers/*.py). */
int ctts_gemm_wgrad_rowmajor(int n_planes, const void* const* dz_planes, const void* const* x_planes, int B, int T, int Cin,
                             int N, int taps, float alpha, int accumulate, float* dw_packed, void* stream);

/* batched plane GEMM with explicit operand views (attention backward products); see ctts_gemm_tc.cu */
int ctts_gemm_batched_planes(int n_planes, const void* const* a_planes, const long long* a_view,
                             const void* const* w_planes, const long long* w_view, const int* addr, long long y_outer,
                             long long y_inner, float alpha, const float* residual, const int64_t* lens, int Z, int T, int K,
                             int N, float* y, void* const* y_planes, void* stream);

/* AlignmentEncoder backward, score part (modules.py:1198-1212): da [B,M,S] = gradient w.r.t. -temp*|q-k|^2 from the
 * gradients of attn_soft / attn_logprob (either may be NULL) */
int ctts_aligner_attention_bwd(const float* soft, const float* logprob, const float* prior, const float* dsoft,
                               const float* dlogprob, const int64_t* src_lens, int B, int M, int S, float* da, void* stream);

/* block-specific pieces: GLU backward (blocks.py:123-134); depthwise Conv1d forward / backward (conformer.py:522-560; the
 * training path needs the un-fused conv because BatchNorm uses batch statistics); relative-shift backward
 * (conformer.py:423-431); fastformer pooling backward (fastformer.py:308-336); elementwise product backward */
int ctts_glu_bwd(const float* h, const float* dg, int rows, int C, float* dh, void* stream);
int ctts_dwconv(const float* x, const float* w, int K, int B, int T, int C, float* y, void* stream);
int ctts_dwconv_bwd(const float* dy, const float* x, const float* w, int K, int B, int T, int C, float* dx, float* dw,
                    void* stream);
int ctts_relshift_bwd(const float* dscore, int Z, int T, int ld, int ld_out, float sqrt_dim, float* dcontent, float* dpos,
                      void* stream);   /* dscore rows have stride ld, dcontent / dpos rows stride ld_out (zero padded) */
int ctts_fastformer_pool_bwd(const float* logits, const float* values, const int64_t* lens, const float* dpooled, int B, int T,
                             int heads, int head_size, float* dlogits, float* dvalues, void* stream);
int ctts_mul_bwd(const float* dy, const float* a, const float* b, int b_rowwise, const int64_t* lens, int B, int T, int C,
                 float* da, float* db, void* stream);

/* liu2021 reference encoder, training only (modules.py:332-397 ReferenceEncoder, coordconv.py:36-71,140-159): channels-last
 * activations [N, H, W, C]; AddCoords(rank 2, with_r) -> [N, H, W, 4]; im2col / col2im of the 3x3, stride (1, 2), pad (1, 1)
 * convolutions (the contraction itself is the dense GEMM engine; column order (kh*3 + kw)*C + c); [R, A, B] -> [R, B, A] */
int ctts_add_coords(const float* x, int N, int H, int W, float* y, void* stream);
int ctts_im2col_3x3_s12(const float* x, int N, int H, int W, int C, float* col, void* stream);
int ctts_col2im_3x3_s12(const float* dcol, int N, int H, int W, int C, float* dx, void* stream);
int ctts_permute_last2(const float* x, int rows, int A, int Bd, float* y, void* stream);

/* single-direction GRU backward through time (liu2021, modules.py:620-640 / :356-392): dgi, dgh [B,T,3H]; dW_hh and db_hh
 * follow by ctts_gemm_generic / ctts_act_bwd on dgh */
int ctts_gru_bwd(const float* gi, const float* w_hh, const float* b_hh, const float* out, int out_ld, int out_off,
                 const float* dout, const float* dh_final, int dhf_ld, int B, int T, int H, int reverse, float* dgi, float* dgh,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTTS_B200_H */
