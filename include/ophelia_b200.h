/* libophelia_sm100.so -- C ABI of the B200-native dc_tts hot path (Text2Mel + SSRN).
 *
 * The reference (CSTR-Edinburgh/ophelia) has no FFI/plugin API: its operator boundary is the Python
 * surface of modules.py / networks.py / architectures.py on top of TensorFlow-1.12 ops.  Each entry point
 * below replaces the TF ops behind one of those functions; the Python mirror in ophelia_b200/ binds them
 * with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (the library never allocates or frees);
 *   - activations are fp32, channels-last [B, time, C] with an explicit row stride `ld*` in elements
 *     (multiple of 4, rows 16-byte aligned);
 *   - conv kernels keep the reference layouts: [k, Cin, Cout] (tf.layers.conv1d) and [1,3,Cout,Cin]
 *     (tf.layers.conv2d_transpose);
 *   - all calls are asynchronous on `stream`, never synchronise, and are legal under CUDA-graph capture;
 *   - return 0 on success, a negative OPH_E* code otherwise; oph_last_error() returns a thread-local message.
 */
#ifndef OPHELIA_B200_H
#define OPHELIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* oph_stream_t; /* cudaStream_t */

#define OPH_OK 0
#define OPH_EINVAL (-1)
#define OPH_ECUDA (-2)

#define OPH_ACT_NONE 0
#define OPH_ACT_RELU 1

#define OPH_PAD_SAME 0
#define OPH_PAD_CAUSAL 1

int oph_version(void);
const char* oph_last_error(void);
/* Host helper (no GPU work): CRC-32C of a byte range, continuing from `crc` (0 to start) -- the checksum of the TF-V2
 * checkpoint format (tensor_bundle records and leveldb table blocks) that ophelia_b200/tf_checkpoint.py reads and writes
 * in place of tf.train.Saver (train.py:296-297, synthesize.py:302-330). */
unsigned int oph_crc32c(const void* data, unsigned long long n, unsigned int crc);

/* ---- instrumentation used by bench.py ---------------------------------------------------------------------
 * oph_launch_count: kernels launched by this library so far (all streams).
 * oph_profile_begin/end: time every launch of the tcgen05 GEMM core with CUDA events on its own stream;
 * oph_profile_begin/end also time the row-wise LayerNorm / highway kernels;
 * out is double[OPH_NUM_TAGS*3] = per tag (launches, summed milliseconds, summed algorithmic FLOPs or bytes). */
#define OPH_TAG_OTHER 0
#define OPH_TAG_CONV_FWD 1
#define OPH_TAG_DGRAD 2
#define OPH_TAG_WGRAD 3
#define OPH_TAG_ATTENTION 4
#define OPH_TAG_ROW_FWD 5   /* LayerNorm / highway forward tails: third value = algorithmic BYTES */
#define OPH_TAG_ROW_BWD 6
#define OPH_TAG_HC_FWD 7      /* the forward GEMM of a highway-conv layer (conv tag 1 = the other conv layers) */
#define OPH_TAG_HC_ROW_FWD 8  /* the highway tail of oph_hc_fwd when it runs as its own launch: BYTES */
#define OPH_TAG_AR_ENC 9      /* oph_ar_encoder_step: third value = fp32 weight BYTES streamed per launch */
#define OPH_NUM_TAGS 10
long long oph_launch_count(void);
/* diagnostics: device buffer long long[74][24]; every GEMM launch overwrites per CTA pair
 * (total cycles, cycles waiting for an accumulator stage, for A, for B, k-blocks). NULL disables. */
int oph_gemm_debug_buffer(long long* dev_buf);
/* diagnostics: a ring of `slots` such buffers; every GEMM launch (also inside a stream capture) takes the next slot, so one
 * replay of a captured step leaves the in-kernel time stamps of all its GEMM launches; _desc = shape of the launch in a slot */
int oph_gemm_debug_ring(long long* dev_buf, int slots);
int oph_gemm_debug_ring_desc(int slot, char* out, int cap);
/* diagnostics (results become garbage): 1 = operand producers skip loads/stores, 2 = weight loader skips copies,
 * 4 = CTA pairs do not rotate their k-block order, 8 = do not feed pre-split operands with TMA tensor copies (producer warps copy them instead),
 * 16 = no remainder K-split of RED outputs, 32 = skip the item-boundary fix-up of flat A tiles,
 * 64 = launch WITH programmatic stream serialization (PDL; off by default: no measured gain), 128 = A copies skipped; bits 8..10: ring depth of the highway-backward row
 * kernel (0 = automatic), 2048 = one block per SM for it, 4096 = previous (warp-per-row) highway-forward row kernel,
 * 32768 = previous (warp-per-row) conv-tail backward kernel, 65536 = fuse the conv tail (LayerNorm / ReLU / dropout / planes) of layers with <= 256 output channels into the
 * GEMM epilogue (off by default: no gain on the training step, slower on the small autoregressive step),
 * 16384 = remainder K-split also for plain conv outputs (memset + RED: forward results then depend on RED order),
 * 131072 = highway tail of oph_hc_fwd always as its own launch (default: inside the conv launch when the row tiles fill whole
 * rounds of the 74 CTA pairs, the last one to >= 70 %), 262144 = also fuse the full rounds when the last round is emptier
 * (the rest of the rows then gets plain work units + a partial tail launch), 524288 = networks.Attention forward as three
 * launches (two GEMMs + the softmax kernel) instead of the one-kernel form, 1048576 = scalar conv-tail kernels for channel
 * counts other than 256 / 512 / 1024 (default: the 16-byte kernels over padded rows), 4194304 = two instead of three blocks per SM
 * for the conv-tail backward kernel */
int oph_gemm_debug_flags(int flags);
/* Every in-kernel barrier wait is bounded and traps instead of hanging the GPU; before it traps the waiting thread leaves a
 * record (block, warp = role, barrier, parity) in mapped host memory.  Returns 1 if such a record exists; out[4] = raw
 * words, text = decoded sentence.  The same sentence is appended to oph_last_error() of the call that sees the failure. */
int oph_last_trap(unsigned long long* out, char* text, int cap);
/* diagnostics: records a fake wait and traps (kills the CUDA context of the calling process): self-test of the record */
int oph_debug_trap_selftest(oph_stream_t stream);
/* Execution context of the calling host thread (like cublasSetStream): with enable != 0 the weight-gradient GEMMs of
 * oph_*_bwd are launched on `side` after an event fork behind the layer's row-wise backward kernel, so that they
 * overlap the input-gradient chain on `stream`.  The caller joins `side` back (event / stream wait) before it reads
 * the gradients and keeps x and dz alive until then.  enable == 0 restores in-order execution. */
int oph_wgrad_stream(oph_stream_t side, int enable);
/* diagnostics: cudaDeviceSetCacheConfig (0 = none, 1 = prefer shared, 2 = prefer L1, 3 = equal) for the calling thread's device */
int oph_cache_config(int mode);
int oph_profile_begin(void);
int oph_profile_end(double* out);

/* ---- weight packing: fp32 kernels -> split-bf16 (hi,lo) SWIZZLE_128B shared-memory images ----------------
 * `deconv`=0: w is [k][Cin][Cout] (modules.py:134-136).  `deconv`=1: w is [3][Cout][Cin] (modules.py:243-250).
 * fwd image feeds oph_*_fwd, bwd image feeds the input-gradient GEMM of oph_*_bwd.
 * Re-pack after every optimiser step (weights changed) or once for inference. */
size_t oph_conv_pack_bytes(int k, int Cin, int Cout, int deconv, int backward);
int oph_conv_pack(const float* w, int k, int Cin, int Cout, int deconv, void* packed_fwd, void* packed_bwd,
                  oph_stream_t stream);
/* Batched form: build a plan on the HOST (array of opaque jobs, oph_pack_job_bytes() each; *njobs / *nblocks are
 * running totals starting at 0), copy it to the device once, then oph_pack_run re-packs every kernel of the plan
 * in ONE launch (called after each optimiser step: the reference re-reads its variables every step for free,
 * architectures.py:128). */
size_t oph_pack_job_bytes(void);
int oph_pack_plan_add(void* plan_host, int capacity, int* njobs, long long* nblocks, const float* w, int k, int Cin,
                      int Cout, int deconv, void* packed_fwd, void* packed_bwd);
int oph_pack_run(const void* plan_dev, int njobs, long long nblocks, oph_stream_t stream);

/* An activation [B*L rows][C]: an fp32 view and/or the same values as split-bf16 planes (hi = bf16(v),
 * lo = bf16(v - hi), row stride ldp elements, ldp % 8 == 0).  Planes are the operand format of the GEMM producers:
 * a layer that receives them copies instead of converting, a layer asked for them (y->hi != NULL, only for
 * C in {256, 512, 1024}) writes them next to the fp32 output.  Unused members are NULL. */
typedef struct {
    float* f32;
    long long ld;
    unsigned short* hi;
    unsigned short* lo;
    long long ldp;
} oph_act;

/* Attention targets that come with the batch instead of the analytic global guide (hp.attention_guide_dir,
 * architectures.py:57-58, 258-280; data_load.py:454-464): w[b][n][t] over the batch-padded [Ng][Tg] block (row stride ld,
 * item stride item_stride, both in floats), `pad` outside that block (1.0 for the guided-attention loss, the value the
 * reference pads the guides with at architectures.py:263; 0.0 for the MSE variant, :275).  mse != 0 selects
 * sum (A - W)^2 (hp.attention_guide_fa) instead of sum |A * W|.  w == NULL keeps the analytic global guide. */
typedef struct {
    const float* w;
    long long item_stride;
    long long ld;
    int Ng;
    int Tg;
    float pad;
    int mse;
    /* gradient inputs of the CDP / Ain / Aout terms (nullable, backward only; written by oph_attention_extra_fwd) */
    const float* col_g;
    const float* col_h;
    float c_aout;
} oph_guide;

/* ---- modules.conv1d (modules.py:91-146), hot-path uses are k=1 -------------------------------------------
 * y = dropout(act(LN(conv(x) + bias))).  z [B*L][ldz] receives the pre-LN conv output (saved for backward,
 * scratch otherwise), stats [B*L][2] = (mean, rstd) (nullable in inference), y_sig (nullable) = sigmoid(LN(..)).
 * in_shift: extra time shift applied to the input rows (AudioEnc C_1 reads mels delayed by one frame,
 * architectures.py:191).  norm: 1 = layer norm (eps 1e-12), 0 = none.  step (nullable, device int64) is mixed
 * into the dropout seed so that a captured graph draws a fresh mask every replay. */
int oph_conv1d_fwd(const oph_act* x, const void* packed_w, const float* bias, const float* gamma, const float* beta,
                   float* z, long long ldz, float* stats, const oph_act* y, float* y_sig, long long ldys, int B, int L,
                   int Cin, int Cout, int k, int rate, int padding, int in_shift, int act, int norm, float drop_p,
                   uint64_t seed, const long long* step, oph_stream_t stream);

/* Backward of oph_conv1d_fwd.  dz [B*L][lddz] is scratch (>= Cout wide; it is rewritten as split-bf16 planes
 * when Cout is 256 or 512).  dx may be NULL (first layer); dw/dbias/dgamma/dbeta are ACCUMULATED into (the
 * caller zeroes the flat gradient buffer once per step). */
int oph_conv1d_bwd(const float* dy, long long lddy, const oph_act* x, const float* z, long long ldz,
                   const float* stats, const void* packed_w_bwd, const float* gamma, const float* beta, float* dz,
                   long long lddz, float* dx, long long lddx, float* dw, float* dbias, float* dgamma, float* dbeta,
                   int B, int L, int Cin, int Cout, int k, int rate, int padding, int in_shift, int act, int norm,
                   float drop_p, uint64_t seed, const long long* step, oph_stream_t stream);

/* ---- modules.normalize (modules.py:47-75) on its own: layer norm over the channels (eps 1e-12, biased variance) --
 * y = LN(x) * gamma + beta for `rows` rows of C channels; stats [rows][2] = (mean, rstd) (nullable; needed by bwd);
 * y's planes are written when given.  Inside conv1d / hc / conv1d_transpose the same arithmetic is fused into the
 * layer's tail; this entry point serves direct calls of the reference function. */
int oph_normalize_fwd(const float* x, long long ldx, const float* gamma, const float* beta, const oph_act* y,
                      float* stats, long long rows, int C, oph_stream_t stream);
/* dx = gradient w.r.t. x (fp32, overwritten); dgamma / dbeta are ACCUMULATED into. */
int oph_normalize_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* stats,
                      const float* gamma, const float* beta, float* dx, long long lddx, float* dgamma, float* dbeta,
                      long long rows, int C, oph_stream_t stream);

/* ---- modules.hc (modules.py:148-207): highway conv, C -> 2C -> C ------------------------------------------
 * z [B*L][ldz] (>= 2C wide) pre-LN conv output; stats [B*L][4] = (mean1, rstd1, mean2, rstd2).
 * x needs its fp32 view (highway residual); its planes, when present, feed the GEMM.
 * One launch when C is 256 / 512 / 1024, x carries planes and the row tiles fill the CTA pairs (see flag 131072): the
 * conv's epilogue warps then apply LN(H1), LN(H2), the sigmoid gate, the highway mix and dropout to their rows while the
 * tensor pipe is on the next row tile.  Otherwise two launches (conv, then the streaming tail kernel). */
int oph_hc_fwd(const oph_act* x, const void* packed_w, const float* bias, const float* g1, const float* b1,
               const float* g2, const float* b2, float* z, long long ldz, float* stats, const oph_act* y, int B, int L,
               int C, int k, int rate, int padding, int norm, float drop_p, uint64_t seed, const long long* step,
               oph_stream_t stream);

/* dz [B*L][lddz] (>= 2C) and dxres [B*L][ldxr] (>= C) are scratch. */
int oph_hc_bwd(const float* dy, long long lddy, const oph_act* x, const float* z, long long ldz,
               const float* stats, const void* packed_w_bwd, const float* g1, const float* b1, const float* g2,
               const float* b2, float* dz, long long lddz, float* dxres, long long ldxr, float* dx, long long lddx,
               float* dw, float* dbias, float* dg1, float* db1, float* dg2, float* db2, int B, int L, int C, int k,
               int rate, int padding, int norm, float drop_p, uint64_t seed, const long long* step,
               oph_stream_t stream);

/* ---- modules.learn_channel_contributions (modules.py:78-88): per-speaker channel gates ---------------------------------
 * gate[b][c] = sigmoid(table[codes[b]][c]) with the zero-padded embedding table [ncodes][C] (code 0 reads as zeros -> 0.5).
 * Behind a conv1d layer (modules.py:141-144): lcc_fwd scales its output y0 (out may alias nothing; out_sig = sigmoid(out),
 * nullable); lcc_bwd gives dy0 = gate * dy and fills scratch [B*L][C] with the per-element table gradients, which
 * lcc_reduce sums over time into dtable (accumulated into; the zero-pad row gets nothing).
 * Inside a highway layer the gate multiplies LN(H2) before the mix (modules.py:200-201): oph_lcc_context arms the NEXT
 * oph_hc_fwd / oph_hc_bwd call of this host thread with (table, codes, scratch); oph_hc_bwd then fills scratch for
 * lcc_reduce.  Gated highway layers run the generic (one warp per row) tail kernels as separate launches. */
int oph_lcc_context(const float* table, const int32_t* codes, float* scratch);
int oph_lcc_fwd(const float* y0, long long ld0, const float* table, const int32_t* codes, const oph_act* out, float* out_sig,
                long long lds, int B, int L, int C, oph_stream_t stream);
int oph_lcc_bwd(const float* dy, long long lddy, const float* y0, long long ld0, const float* table, const int32_t* codes,
                float* dy0, long long ldd0, float* scratch, int B, int L, int C, oph_stream_t stream);
int oph_lcc_reduce(const float* scratch, const int32_t* codes, float* dtable, int B, int L, int C, oph_stream_t stream);

/* ---- modules.conv1d_transpose (modules.py:209-258): stride-2, k=3, always layer-normed ---------------------
 * out[2i] = W0.x[i] + W2.x[i-1], out[2i+1] = W1.x[i]; y is [B][2L][C].  z [B*2L][ldz], stats [B*2L][2]. */
int oph_deconv_fwd(const oph_act* x, const void* packed_w, const float* bias, const float* gamma, const float* beta,
                   float* z, long long ldz, float* stats, const oph_act* y, int B, int L, int C, float drop_p,
                   uint64_t seed, const long long* step, oph_stream_t stream);
int oph_deconv_bwd(const float* dy, long long lddy, const oph_act* x, const float* z, long long ldz,
                   const float* stats, const void* packed_w_bwd, const float* gamma, const float* beta, float* dz,
                   long long lddz, float* dx, long long lddx, float* dw, float* dbias, float* dgamma, float* dbeta,
                   int B, int L, int C, float drop_p, uint64_t seed, const long long* step, oph_stream_t stream);

/* ---- modules.embed (modules.py:15-44) -------------------------------------------------------------------- */
int oph_embed_fwd(const int32_t* ids, const float* table, float* out, long long ldo, int rows, int E,
                  oph_stream_t stream);
int oph_embed_bwd(const int32_t* ids, const float* dout, long long ldo, float* dtable, int rows, int E,
                  oph_stream_t stream);

/* ---- networks.Attention (networks.py:286-325) --------------------------------------------------------------
 * A = softmax(Q K^T / sqrt(d)) [B][T][ldA] (ldA >= N), R = A V written with row stride ldr (so it can land in
 * the first half of the [R, Q] buffer, networks.py:317-319).  prev_max (nullable, int32 [B]) + win enable the
 * forcibly-incremental window: keys outside [prev, prev+win) get -2^32+1 (networks.py:304-313); with win <= 0 prev_max
 * holds a per-item key count instead and keys n >= prev_max[b] are masked (hp.turn_off_monotonic_for_synthesis,
 * networks.py:307-309).
 * align_t (nullable) = alignments [B][N][T]; argmax (nullable) int32 [B][T] (first maximum);
 * att_acc (nullable, device double) += sum A*W over n<maxN, t<maxT with the analytic guide (utils.py:155-161), or with
 * the batch's own targets when guide != NULL (see oph_guide). */
/* Q, K, V, A (probabilities) and R (context vectors) are activations per batch item; for the outputs A and R the f32
 * view is written and the planes too when given (R's by the epilogue of the A.V product, so that the decoder's first
 * conv reads [R|Q] through the copy engines).  When every
 * operand carries split-bf16 planes (contiguous items: item z starts z*rows*ldp elements in), both products are fed by
 * the copy engines; otherwise the producer warps convert the fp32 views. */
int oph_attention_fwd(const oph_act* Q, const oph_act* K, const oph_act* V, const oph_act* A, const oph_act* R,
                      float* align_t, int32_t* argmax, const int32_t* prev_max, int win, double* att_acc, int maxN,
                      int maxT, float g, int B, int T, int N, int d, const oph_guide* guide, oph_stream_t stream);
/* dR [B][T].  dA [B][T][ldA] scratch (f32 + optional planes for dS).  dq_addend (nullable) is added into dQ (the direct
 * [R,Q] concat path).  att_coef = lw_att / (B*min(N,maxN)*min(T,maxT)) injects the guided-attention gradient. */
int oph_attention_bwd(const oph_act* dR, const oph_act* Q, const oph_act* K, const oph_act* V, const oph_act* A,
                      const oph_act* dA, float* dQ, long long lddq, const float* dq_addend, long long ldqa, float* dK,
                      long long lddk, float* dV, long long lddv, float att_coef, int maxN, int maxT, float g, int B,
                      int T, int N, int d, const oph_guide* guide, oph_stream_t stream);
/* networks.FixedAttention (networks.py:327-358) takes its alignments A [B][T][ldA] from outside (external durations) but
 * reports the same guided-attention term: att_acc += sum A*W over n < maxN, t < maxT (guide as in oph_attention_fwd). */
int oph_attention_guide_sum(const float* A, long long ldA, int B, int T, int N, double* att_acc, int maxN, int maxT, float g,
                            const oph_guide* guide, oph_stream_t stream);
/* "Confidence through attention" losses (hp.lw_cdp / lw_ain / lw_aout, architectures.py:283-321) over the alignments
 * A [B][T][ldA] of a training batch: coverage deviation penalty and the two attention entropies.
 * extra_fwd: acc3 (device double[3], zeroed by the caller) += (sum log(1 + (1 - s)^2), sum_t P log P summed over keys,
 *   sum A log A);  col_g / col_h [B][N] receive the per-key factors of the gradient for oph_attention_bwd
 *   (oph_guide.col_g / col_h), with c_cdp = lw_cdp / (B N), c_ain = -lw_ain / (B N log T); oph_guide.c_aout =
 *   -lw_aout / (B T log N).
 * extra_finalize: comps8[5..7] = CDP, Ain, Aout; add_to_total != 0 adds the weighted terms to comps8[0] (the legacy lw_*
 *   pattern of architectures.py:333-349; the loss_weights dict pattern reports them without adding them). */
int oph_attention_extra_fwd(const float* A, long long ldA, int B, int T, int N, float c_cdp, float c_ain, float* col_g,
                            float* col_h, double* acc3, oph_stream_t stream);
int oph_attention_extra_finalize(const double* acc3, float* comps8, int B, int T, int N, float w_cdp, float w_ain,
                                 float w_aout, int add_to_total, oph_stream_t stream);
/* fp32 rows -> split-bf16 planes for operands that do not come out of a row-wise kernel of this library. */
int oph_split_planes(const float* x, long long ldx, long long rows, int C, unsigned short* hi, unsigned short* lo,
                     long long ldp, oph_stream_t stream);

/* ---- losses (architectures.py:147-173, 245-355) ------------------------------------------------------------
 * acc: device double[4] zeroed by the caller: sum|Y-t|, sum BCE, sum (Y-t)^2, (attention sum).
 * dlogits (nullable) receives d(loss)/d(logits) for the weighted sum of the three reconstruction terms. */
int oph_recon_loss(const float* logits, long long ldl, const float* target, long long ldt, float* dlogits,
                   long long ldd, long long rows, int C, int squash, float w_l1, float w_bd, float w_l2,
                   double* acc, oph_stream_t stream);
int oph_loss_finalize(const double* acc, float* out, double n_recon, double n_att, float w_l1, float w_bd,
                      float w_att, float w_l2, int has_att, int squash, oph_stream_t stream);

/* ---- optimiser (architectures.py:96-131, utils.py:167-170) ------------------------------------------------
 * lr_t: device float[2] = (Adam step size incl. bias correction, scheduled lr). */
int oph_adam_prepare(const long long* global_step, float* lr_t, float lr0, float beta1, float beta2, int decay_lr,
                     float warmup, oph_stream_t stream);
int oph_adam_clip(float* p, float* m, float* v, const float* g, long long n, const float* lr_t, float beta1,
                  float beta2, float eps, float clip, float grad_scale, oph_stream_t stream);
int oph_step_inc(long long* global_step, oph_stream_t stream);

/* ---- raw access to the tcgen05 GEMM core (tests / diagnostics) --------------------------------------------
 * C[M][N] = alpha * A[M][K] * op(B) (+ bias) with 3-term split-bf16; b_mode 1: B is [N][K]; 2: B is [K][N]. */
int oph_gemm_nt(const float* A, long long lda, const float* Bm, long long ldb, float* C, long long ldc,
                const float* bias, int M, int N, int K, int b_mode, float alpha, int batch, long long a_bs,
                long long b_bs, long long c_bs, oph_stream_t stream);
/* C[M][N] (+)= A^T B with A [R][lda] (M columns), B [R][ldb] (N columns): the weight-gradient form. */
int oph_gemm_tn(const float* A, long long lda, const float* Bm, long long ldb, float* C, long long ldc, int M,
                int N, int R, int splits, oph_stream_t stream);

/* ---- autoregressive frame step, incremental (synthesize.py:150-230) ---------------------------------------
 * The reference recomputes AudioEnc + Attention + AudioDec over all max_T frames per generated frame and keeps row j.
 * AudioEnc is causal and its input row t is final once frame t - 1 exists, so row j of a layer needs rows j, j - rate,
 * j - 2 rate of the layer below only: every layer keeps its output history [B][T][ld] (item stride *_item floats) and
 * conv_step / hc_step compute row j = *frame (device int, so one captured CUDA graph serves all frames) for all
 * B <= 16 items.  fp32 FMA arithmetic, deterministic.
 * w is the reference-layout kernel [k][Cin][Cout] (fp32, not packed).  gamma == NULL: no layer norm (hp.norm = None).
 * scratch: device floats, at least oph_ar_scratch_floats().
 *  conv_step: y[b][j] = act(LN(bias + sum_i W[i] x[b][j - in_shift - (k-1-i) rate]));  y_sig (nullable) = sigmoid(LN(..))
 *  hc_step:   highway layer (modules.py:148-207) with [H1 | H2] = conv to 2C channels and the residual x[b][j]
 *  window_gather / window_scatter: Attention cannot be cached -- the window mask of the latest prev_max_attentions
 *             applies to every time row (networks.py:304-313), so R[t < j] changes when the window moves and AudioDec's
 *             row j sees it through its causal reach.  oph_attention_fwd and the AudioDec layers are therefore run per
 *             step over the W = reach + 1 rows [s, s + W), s = max(0, min(j - reach, T - W)): gather copies those rows
 *             of the Q history into a static [B][W][ldw] buffer, scatter moves row j - s of the window results
 *             (Yw [B][W][ldyw], align_w [B][N][W], argmax_w [B][W]) into frame j of Y [B][T][ldy], alignments
 *             [B][N][T], prev [B] and history [T][B] (synthesize.py:204-209).
 *  advance:   *frame += 1 */
/* One layer of the AudioEnc frame step for oph_ar_encoder_step: the same quantities as the arguments of conv_step / hc_step
 * (kind 0: conv1d with C = Cout output channels; kind 1: highway layer with C channels, w = [k][C][2C]; g2 / b2 for
 * highway layers only; g1 == NULL: no layer norm). */
typedef struct {
    const float* w;
    const float* bias;
    const float* g1;
    const float* b1;
    const float* g2;
    const float* b2;
    const float* x;
    float* y;
    long long x_item, ldx, y_item, ldy;
    int Cin, C, k, rate, kind, act, in_shift;
} oph_ar_layer;
/* The whole AudioEnc frame step in one launch: row j of an item depends on that item's histories only, so the layers of an
 * item run back to back in one kernel; an item gets a thread-block cluster of 8 CTAs (CTA r owns 1/8 of every layer's
 * output channels, LayerNorm moments through distributed shared memory, a cluster barrier between layers).  Same results
 * as the per-layer calls up to the summation order.  Needs C / 8 (2 C / 8 for highway layers) in {32, 64, 128}. */
int oph_ar_encoder_step(const oph_ar_layer* layers, int nlayers, int B, const int* frame, oph_stream_t stream);
size_t oph_ar_scratch_floats(void);
int oph_ar_conv_step(const float* x, long long x_item, long long ldx, const float* w, const float* bias,
                     const float* gamma, const float* beta, float* y, long long y_item, long long ldy, float* y_sig,
                     long long s_item, long long lds, float* scratch, int B, int Cin, int Cout, int k, int rate,
                     int in_shift, int act, const int* frame, oph_stream_t stream);
int oph_ar_hc_step(const float* x, long long x_item, long long ldx, const float* w, const float* bias, const float* g1,
                   const float* b1, const float* g2, const float* b2, float* y, long long y_item, long long ldy,
                   float* scratch, int B, int C, int k, int rate, const int* frame, oph_stream_t stream);
int oph_ar_window_gather(const float* Q, long long q_item, long long ldq, float* Qw, long long w_item, long long ldw,
                         int B, int d, int T, int W, int reach, const int* frame, oph_stream_t stream);
int oph_ar_window_scatter(const float* Yw, long long yw_item, long long ldyw, float* Y, long long y_item, long long ldy,
                          int n_mels, const float* align_w, float* align_t, const int32_t* argmax_w, int32_t* prev,
                          int32_t* history, int B, int N, int T, int W, int reach, const int* frame,
                          oph_stream_t stream);
int oph_ar_advance(int32_t* frame, oph_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
