/* dpc_b200.h -- C ABI of libdpc_b200.so: the B200 (sm_100a) kernels behind the DPC-RNN training path.
 *
 * The reference (TengdaHan/DPC) has no FFI: its hot path is a chain of stock ATen operators called
 * from Python nn.Modules.  Each entry point below replaces the operator call sites cited next to
 * it (paths relative to /root/reference).  The library is loaded with ctypes (dpc_b200/_lib.py);
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching allocator);
 *     the library never allocates or frees on the hot path and keeps no pointer after returning;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 = ok, non-zero = error; dpc_last_error() returns a thread-local message;
 *   - activations are channels-last rows: a tensor [NB,T,H,W,C] fp32 is `rows = NB*T*H*W` rows of C;
 *   - no C++ exceptions, no ATen / pybind types cross this boundary.
 */
#ifndef DPC_B200_H
#define DPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPC_B200_ABI_VERSION 1

/* Geometry of one Conv3d call site.  in [NB,Ti,Hi,Wi,Ci] -> out [NB,To,Ho,Wo,Co]. */
typedef struct dpc_conv_geom {
    int32_t NB, Ti, Hi, Wi, Ci;
    int32_t To, Ho, Wo, Co;
    int32_t kT, kH, kW;
    int32_t sT, sH, sW;
    int32_t pT, pH, pW;
} dpc_conv_geom;

/* ---- library ---------------------------------------------------------------------------- */
int dpc_abi_version(void);
const char* dpc_last_error(void);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
int64_t dpc_launch_count(void);

/* ---- tcgen05 tensor-core path (3xBF16 split: fp32-equivalent to ~1e-5) ----------------------
 * Operands are pairs of bf16 planes: hi = bf16(x), lo = bf16(x - hi), multiplied as hi*hi + hi*lo + lo*hi into fp32 TMEM
 * accumulators.  Replaces nn.Conv3d forward / backward at backbone/resnet_2d3d.py:13-31 (conv3x3x3, conv1x3x3), :241-244
 * (downsample 1x1x1) as used by BasicBlock2d/3d.forward (:64-80, :100-116) and Bottleneck2d/3d.forward, plus torch.matmul at
 * dpc/model_3d.py:83.  Channel counts must be multiples of 64.  tcgen05 kind::f16 needs ONE input format per
 * instruction, and wgrad multiplies activations by gradients (which need bf16's exponent range), so activations are bf16
 * pairs too; a GEMM whose operands are both forward values may use fp16 pairs (dpc_split_f16, ~22 mantissa bits). */
int dpc_split_bf16(const float* src, void* hi, void* lo, int64_t n, void* stream);
int dpc_split_f16(const float* src, void* hi, void* lo, int64_t n, void* stream);
/* w [Co,Ci,taps] -> forward planes wf_* [Co][tap][Ci] and dgrad planes wd_* [Ci][tap][Co];
 * either pair may be NULL */
int dpc_pack_conv_weight_bf16(const float* w, void* wf_hi, void* wf_lo, void* wd_hi, void* wd_lo,
                              int Co, int Ci, int taps, void* stream);
/* C[M,N] (+)= A[M,K] * B[N,K]^T, fp32 out, K % 64 == 0; f16: both operands are fp16 pairs (else bf16 pairs) */
int dpc_gemm_nt_split_tc(int M, int N, int K, const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                         int f16, float* C, int accumulate, void* stream);
/* the score matmul of dpc/model_3d.py:83 on a persistent, A-resident schedule (score_tc.cu): C[M,N] = A[M,256] . B[N,256]^T,
 * K must be 256; f16 as above */
int dpc_score_matmul_tc(int M, int N, int K, const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int f16,
                        float* C, void* stream);
/* forward, strides in {1,2}: y = conv(x planes [NB,Ti,Hi,Wi,Ci], wf planes [Co][taps][Ci]).
 * bn_ws (nullable, 2*Co doubles): the epilogue also accumulates the per-channel sum / sum of squares of y
 * (the BatchNorm batch statistics), to be turned into mean / rstd by dpc_bn_finalize. */
int dpc_conv3d_fwd_tc(const dpc_conv_geom* g, const void* x_hi, const void* x_lo, const void* wf_hi,
                      const void* wf_lo, float* y, double* bn_ws, void* stream);
/* dgrad: dx (+)= conv^T(dy planes [NB,To,Ho,Wo,Co], wd planes [Ci][taps][Co]); one launch per
 * input-parity class of a strided site */
int dpc_conv3d_dgrad_tc(const dpc_conv_geom* g, const void* dy_hi, const void* dy_lo, const void* wd_hi,
                        const void* wd_lo, float* dx, int accumulate, void* stream);
/* dgrad of a STRIDE-1 site with the BatchNorm-backward reduction of the consumer BN fused into the epilogue
 * (resnet_2d3d.py:55-78 backward): with v = the final dx, g = v * [mask_hi > 0] (mask_hi: hi plane of that BN's
 * ReLU output, nullable) and xhat = (y - mean) * rstd:  ws[0..Ci) = sum_rows g, ws[Ci..2Ci) = sum_rows g * xhat. */
int dpc_conv3d_dgrad_bnred_tc(const dpc_conv_geom* g, const void* dy_hi, const void* dy_lo, const void* wd_hi,
                              const void* wd_lo, float* dx, int accumulate, const void* mask_hi, const float* y,
                              const float* mean, const float* rstd, double* ws, void* stream);
/* wgrad: dw [Co,Ci,kT,kH,kW] = sum over positions dy (x) x; dwp = scratch [Co][taps][Ci] fp32 */
int dpc_conv3d_wgrad_tc(const dpc_conv_geom* g, const void* x_hi, const void* x_lo, const void* dy_hi,
                        const void* dy_lo, float* dwp, float* dw, void* stream);

/* ---- stem: Conv3d(3,64,(1,7,7),s(1,2,2),p(0,3,3)) reading the caller's NCDHW input ----------
 * replaces backbone/resnet_2d3d.py:211,260 (self.conv1).  x [NB,3,T,H,W] -> y [NB,T,H/2,W/2,64].
 * Fallback for odd frame sizes (3xBF16 split on tcgen05; the im2col tile is built in shared memory from the fp32 video).
 * bn_ws (nullable, 128 doubles) receives the per-channel sum | sum of squares of y (bn1 statistics). */
int dpc_stem_conv_fwd_tc(const float* x, const float* w, float* y, double* bn_ws, int NB, int T, int H, int W,
                         void* stream);
/* dw [64,3,1,7,7] from the video and the split-bf16 planes of dy [NB,T,H/2,W/2,64] */
int dpc_stem_conv_wgrad_tc(const float* x, const void* dy_hi, const void* dy_lo, float* dw, int NB, int T,
                           int H, int W, void* stream);
/* Space-to-depth formulation of conv1 (same reference site): the stride-2 7x7 conv is a stride-1 4x4 conv over
 * X2[h2,w2][(c,r,s)] = x[c,2*h2+r,2*w2+s], stored as split-bf16 planes [NB,T,H/2,W/2,16] (12 real channels).
 * dpc_stem_s2d_pack builds the planes (H, W even). */
int dpc_stem_s2d_pack(const float* x, void* x2_hi, void* x2_lo, int NB, int T, int H, int W, void* stream);
/* dw [64,3,1,7,7] from the space-to-depth planes and the split-bf16 planes of dy [NB,T,H/2,W/2,64] */
int dpc_stem_conv_wgrad_s2d(const void* x2_hi, const void* x2_lo, const void* dy_hi, const void* dy_lo, float* dw,
                            int NB, int T, int H, int W, void* stream);

/* Pooled stem (stem_pool.cu): conv1 + bn1 + relu + maxpool (backbone/resnet_2d3d.py:211-214,260-263) without ever storing
 * the conv1 output.  Forward: dpc_stem_s2d_pack -> dpc_stem_s2d_wpack -> dpc_stem_pool_fwd (per pooled position the conv1
 * value max-pool o relu o bn1 selects + its 3x3-window index; bn1 batch sums over all conv positions) -> dpc_bn_finalize ->
 * dpc_stem_pool_finalize (normalised, ReLU'd operand planes of layer1; ReLU-dead windows flagged in idx).  Backward:
 * dpc_stem_pool_bwd_reduce (bn1 backward sums on the pooled grid) -> dpc_stem_pool_bwd (conv1 recomputed; gradient planes on
 * the conv1 grid) -> dpc_stem_conv_wgrad_s2d.  Pooled extents: Hp = (H/2 - 1)/2 + 1, Wp likewise; rows = NB*T*Hp*Wp. */
int dpc_stem_pool_supported(int H, int W);
int dpc_stem_s2d_wpack(const float* w /*[64,3,1,7,7]*/, void* wp /*32768 bf16*/, void* stream);
int dpc_stem_pool_fwd(const void* x2_hi, const void* x2_lo, const void* wp, const float* gamma, float* ypool /*[rows,64]*/,
                      void* idx /*[rows,64] uint8*/, double* bn_ws /*128, nullable*/, int NB, int T, int H, int W, void* stream);
int dpc_stem_pool_finalize(const float* ypool, void* idx, const float* mean, const float* rstd, const float* gamma,
                           const float* beta, void* a_hi, void* a_lo, float* a_rows /*nullable*/, int64_t rows, void* stream);
int dpc_stem_pool_bwd_reduce(const float* ypool, const float* dout, const void* idx, const float* mean, const float* rstd,
                             double* ws /*128*/, float* dgamma, float* dbeta, int64_t rows, void* stream);
int dpc_stem_pool_bwd(const void* x2_hi, const void* x2_lo, const void* wp, const float* dout, const void* idx,
                      const float* mean, const float* rstd, const float* gamma, const double* ws, void* dy_hi, void* dy_lo,
                      int NB, int T, int H, int W, void* stream);
/* same, with conv1's weight gradient dw [64,3,1,7,7] computed by the same kernel (needs dpc_stem_pool_supported(H, W) == 2) */
int dpc_stem_pool_bwd_wgrad(const void* x2_hi, const void* x2_lo, const void* wp, const float* dout, const void* idx,
                            const float* mean, const float* rstd, const float* gamma, const double* ws, float* dw,
                            int NB, int T, int H, int W, void* stream);

/* ---- BatchNorm3d(track_running_stats=False): batch statistics always ----------------------
 * replaces nn.BatchNorm3d at resnet_2d3d.py:55,59,91,95,212,243 (+ relu_ / `out += residual`
 * at :68-78,:104-114).  Biased variance, eps as given (1e-5).  `ws` = 2*C doubles of scratch. */
int dpc_bn_stats(const float* y, int64_t rows, int C, double* ws, float* mean, float* rstd,
                 float eps, void* stream);
int dpc_bn_finalize(const double* ws, int64_t rows, int C, float eps, float* mean, float* rstd, void* stream);
/* out = [relu]( bn(y) + residual ), residual = none | res | bn_r(res)  (downsample branch).
 * The raw residual may be fp32 rows (`res`) or split-bf16 planes (`res_hi/res_lo`); the result is written
 * as fp32 rows (`out`, nullable) and/or split-bf16 planes (`out_hi/out_lo`, nullable) -- the
 * tensor-core operand format, so no separate conversion pass is needed. */
int dpc_bn_apply_fwd(const float* y, const float* mean, const float* rstd, const float* gamma,
                     const float* beta, const float* res, const void* res_hi, const void* res_lo,
                     const float* r_mean, const float* r_rstd, const float* r_gamma, const float* r_beta,
                     int relu, float* out, void* out_hi, void* out_lo, int64_t rows, int C, void* stream);
/* backward of the above for ONE BatchNorm: g = dout * (out > 0 if relu), then
 * dgamma = sum g*xhat, dbeta = sum g, dy = gamma*rstd*(g - dbeta/n - xhat*dgamma/n).
 * The ReLU mask comes from `out` (fp32 rows) or `out_hi` (the hi plane).  dy is written as fp32 rows
 * and/or split-bf16 planes.  g_out (nullable) receives g (the gradient of the identity residual).
 * `ws` = 2*C doubles. */
int dpc_bn_bwd(const float* dout, const float* out, const void* out_hi, int relu, const float* y,
               const float* mean, const float* rstd, const float* gamma, double* ws, float* dgamma,
               float* dbeta, float* dy, void* dy_hi, void* dy_lo, float* g_out, int64_t rows, int C,
               void* stream);
/* The same with the reduction supplied by the caller (ws = [sum g | sum g*xhat], 2*C doubles): finalize + apply only.
 * dpc_conv3d_dgrad_bnred_tc produces such a ws in the epilogue of the dgrad that computes `dout`. */
int dpc_bn_bwd_apply(const float* dout, const float* out, const void* out_hi, int relu, const float* y,
                     const float* mean, const float* rstd, const float* gamma, const double* ws, float* dgamma,
                     float* dbeta, float* dy, void* dy_hi, void* dy_lo, float* g_out, int64_t rows, int C,
                     void* stream);

/* ---- stem tail: BN + ReLU + MaxPool3d((1,3,3),s(1,2,2),p(0,1,1)) in one pass ---------------
 * replaces resnet_2d3d.py:212-214,261-263.  y [NB*T,H,W,C] -> out [NB*T,H/2,W/2,C]. */
int dpc_bn_relu_maxpool_fwd(const float* y, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, float* out, int NT, int H, int W, int C, void* stream);
/* g [NB*T,H,W,C] = gradient w.r.t. bn(y) (ReLU and pool already undone), from dout on the pooled grid */
int dpc_bn_relu_maxpool_bwd(const float* y, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, const float* out, const float* dout, float* g,
                            int NT, int H, int W, int C, void* stream);

/* the whole stem tail backward in two passes (no materialised g): max-pool bwd + ReLU bwd + bn1 bwd.
 * dout lives on the pooled grid; dy (fp32 rows and/or split-bf16 planes) on the conv1 output grid.
 * pooled_reduce != 0: the reduction pass runs on the pooled grid only (xhat of each window's arg-max is
 * recovered from the pooled output: needs gamma != 0 everywhere). */
int dpc_stem_tail_bwd(const float* y, const float* mean, const float* rstd, const float* gamma,
                      const float* beta, const float* out, const float* dout, double* ws, float* dgamma,
                      float* dbeta, float* dy, void* dy_hi, void* dy_lo, int NT, int H, int W, int C,
                      int pooled_reduce, void* stream);

/* ---- temporal average + ReLU split --------------------------------------------------------
 * replaces F.avg_pool3d / self.relu at dpc/model_3d.py:53-57.  z [NB,T,S,C] ->
 * finf [NB,S,C] (mean over T, pre-ReLU) and feat [NB,S,C] (ReLU). */
int dpc_pool_split_fwd(const float* z, float* finf, float* feat, int NB, int T, int S, int C, void* stream);
int dpc_pool_split_bwd(const float* finf, const float* dfinf, const float* dfeat, float* dz,
                       int NB, int T, int S, int C, void* stream);

/* ---- dense helpers used by the ConvGRU / predictor / score stages ---------------------------
 * C[M,N] (ldc) = alpha * opA(A) * opB(B) + beta * C;  row-major; transX != 0 means the operand is
 * stored transposed.  fp32 CUDA-core GEMM: the small (latency-bound) GEMMs of convrnn.py:29-33 and
 * model_3d.py:36-40,68. */
int dpc_gemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                 const float* B, int ldb, float beta, float* C, int ldc, void* stream);
/* dst[r] = src[(r / inner) * outer + offset + r % inner]  (rows of D floats) */
int dpc_gather_rows(const float* src, float* dst, int64_t rows, int D, int64_t inner, int64_t outer,
                    int64_t offset, void* stream);
/* dst[(r / inner) * outer + offset + r % inner] (+)= src[r] */
int dpc_scatter_rows(const float* src, float* dst, int64_t rows, int D, int64_t inner, int64_t outer,
                     int64_t offset, int accumulate, void* stream);
/* colsum[n] (+)= sum_r A[r,n] */
int dpc_colsum(const float* A, int64_t rows, int N, float* out, int accumulate, void* stream);

/* ---- ConvGRU cell (kernel_size 1) gate math ------------------------------------------------
 * replaces torch.sigmoid / tanh / mul / add / Dropout at backbone/convrnn.py:29-33,78.
 * pre-activations come from dpc_gemm_f32; all tensors [R,D] unless noted. */
int dpc_gru_gates_zr(const float* xz, const float* xr, int ldx, const float* hzr /*[R,2D]: z|r*/,
                     const float* bz, const float* br, const float* h, float* z, float* r,
                     float* hr, int64_t R, int D, void* stream);
/* o = tanh(xo + ho + bo); hn = h*(1-z) + o*z; hout = hn * keep(seed,step)/(1-p) (p == 0: no dropout) */
int dpc_gru_out(const float* xo, int ldx, const float* ho, const float* bo, const float* h,
                const float* z, float* o, float* hout, float* keep /*nullable [R,D]*/,
                float p, uint64_t seed, uint64_t offset, int64_t R, int D, void* stream);
/* backward of one cell step, elementwise part (see dpc_b200/engine.py for the GEMMs around it) */
int dpc_gru_bwd_out(const float* dhout, const float* keep, const float* h, const float* z,
                    const float* o, float* dpre_o, float* dpre_z_partial, float* dh, int64_t R, int D,
                    void* stream);
int dpc_gru_bwd_zr(const float* dhr, const float* h, const float* r, const float* z,
                   const float* dz_partial, float* dpre_zr /*[R,2D]*/, float* dh /*accumulated*/,
                   int64_t R, int D, void* stream);
/* y = relu(x + b) (b nullable) and its backward */
int dpc_bias_relu(const float* x, const float* b, float* y, int relu, int64_t R, int D, void* stream);
int dpc_relu_bwd(const float* y, const float* dy, float* dx, int accumulate, int64_t n, void* stream);

/* ---- recurrent head as one kernel per direction (head_chain.cu; feature size D = 256, ConvGRU kernel_size 1) ---------
 * replaces ConvGRUCell.forward / ConvGRU.forward (backbone/convrnn.py:24-34,62-88) over the N - P aggregated blocks and the
 * prediction loop of dpc/model_3d.py:62-72.  Rows r = b*S + s (R = B*S); T7 = N - 1 GRU steps run (N - P aggregate +
 * P - 1 in the prediction loop; the step after the last prediction is dead work).  Gate weights W* are [D][2D] (x | h
 * columns), network_pred weights [D][D].  Saved tensors are step-major:
 *   XH [T7][R][2D] gate input (x | h), XO [T7][R][2D] out-gate input (x | h*r), Z / Rg / O / Keep [T7][R][D],
 *   U / Hp / Pp [P][R][D] network_pred hidden (post-ReLU) / the state it was predicted from / the prediction,
 *   pred_rows [B*P*S][D] the predictions in score-row order.
 * Backward: dfeat [B*N*S][D] (only the aggregated blocks are written), DZR [T7][R][2D], DO [T7][R][D], DP / DU [P][R][D]
 * = gradients of the pre-activations; the weight gradients are reductions over (steps x rows) of those against XH / XO /
 * U / Hp (dpc_conv3d_wgrad_tc) and their column sums (dpc_colsum). */
int dpc_head_chain_pack(const float* Wz, const float* Wr, const float* Wo, const float* W0, const float* W2,
                        float* wt_zr /*[2D][2D]*/, float* wt_o /*[2D][D]*/, float* w0t, float* w2t, void* stream);
int dpc_head_chain_fwd(const float* feat /*[B*N*S][D]*/, const float* wt_zr, const float* wt_o, const float* w0t,
                       const float* w2t, const float* bz, const float* br, const float* bo, const float* b0, const float* b2,
                       int B, int N, int S, int P, float p_drop, uint64_t seed, float* XH, float* XO, float* Z, float* Rg,
                       float* O, float* Keep /*nullable when p_drop == 0*/, float* U, float* Hp, float* Pp, float* pred_rows,
                       void* stream);
int dpc_head_chain_bwd(const float* dpred_rows, const float* Wz, const float* Wr, const float* Wo, const float* W0,
                       const float* W2, int B, int N, int S, int P, const float* XH, const float* Z, const float* Rg,
                       const float* O, const float* Keep, const float* U, const float* Pp, float* dfeat, float* DZR, float* DO,
                       float* DP, float* DU, void* stream);

/* ---- NCE score / mask / cross-entropy -------------------------------------------------------
 * mask: closed form of the Python loops at dpc/model_3d.py:86-96 (values {1,-1,-3,0}, contiguous).
 * ce: nn.CrossEntropyLoss(mean) with target = diagonal, dpc/main.py:178-185,213-217, and
 * calc_topk_accuracy (utils/utils.py:38-55) for k = 1,3,5. */
int dpc_nce_mask_fill(int8_t* mask, int B, int P, int SQ, void* stream);
/* score [rows, M]; positive of row i = column i % M (rows == M on one device; rows = n_gpu*M
 * after the reference's DataParallel gather).  out[0] = loss, out[1..3] = top1/3/5 accuracy;
 * lse[rows] saved for backward */
int dpc_nce_ce_fwd(const float* score, int rows, int M, float* lse, float* out, void* stream);
/* dscore = gscale[0] * (softmax - onehot) / rows   (gscale: device scalar, nullable = 1) */
int dpc_nce_ce_bwd(const float* score, const float* lse, const float* gscale, float* dscore,
                   int rows, int M, void* stream);

/* ---- LC classifier pieces (eval/model_3d_lc.py:12-65) ------------------------
 * BatchNorm with running statistics (track_running_stats=True): train mode normalises with batch statistics
 * and updates the buffers (momentum, unbiased variance); eval mode normalises with the buffers. */
int dpc_bn_running_update(const float* mean, const float* rstd, int64_t rows, float eps, float momentum,
                          float* running_mean, float* running_var, int C, void* stream);
int dpc_bn_rstd_from_var(const float* var, float eps, float* rstd, int C, void* stream);
/* feat[n,e] = mean_t relu(z[n,t,e])  (LC applies ReLU BEFORE the temporal average, model_3d_lc.py:53-55) */
int dpc_relu_pool_fwd(const float* z, float* feat, int NB, int T, int64_t E, void* stream);
int dpc_relu_pool_bwd(const float* z, const float* dfeat, float* dz, int NB, int T, int64_t E, void* stream);
/* y = x * keep, keep in {0, 1/(1-p)} (nn.Dropout before the final Linear, model_3d_lc.py:43) */
int dpc_dropout_fwd(const float* x, float* y, float* keep, float p, uint64_t seed, uint64_t offset, int64_t n,
                    void* stream);
int dpc_mul(const float* a, const float* b, float* out, int64_t n, void* stream);

/* ---- optimiser ------------------------------------------------------------------------------
 * torch.optim.Adam(lr, weight_decay) (L2, not AdamW), dpc/main.py:81,231, over a flat buffer. */
int dpc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, float wd, int step, float gscale, void* stream);

/* ---- data-parallel exchange (comm.cu) -------------------------------------------------------
 * ONE all-reduce (sum, in place) of the flat fp32 gradient buffer over NCCL on the caller's stream: the gradient reduction
 * of nn.DataParallel's backward (dpc/main.py:65,230) in the one-process-per-GPU layout.  NCCL is taken from the libnccl.so.2
 * already loaded into the process.  dpc_comm_unique_id: rank 0 fills 128 bytes that the host distributes to every rank;
 * dpc_comm_init: collective over all ranks (current CUDA device = this rank's GPU) -> opaque communicator. */
#define DPC_COMM_ID_BYTES 128
int dpc_comm_unique_id(void* id128 /*host, DPC_COMM_ID_BYTES*/);
int dpc_comm_init(const void* id128 /*host*/, int rank, int world, void** comm);
int dpc_flat_allreduce(void* comm, float* buf, int64_t n, void* stream);
int dpc_comm_destroy(void* comm);

/* ---- on-device clip augmentation (augment.cu): decoded uint8 frames -> the float32 block DPC_RNN.forward consumes --------
 * replaces the CPU transform chain utils/augmentation.py:147-384 composed as dpc/main.py:115-133 (RandomSizedCrop | RandomCrop +
 * Scale, RandomHorizontalFlip, RandomGray, ColorJitter, ToTensor, Normalize) and the reshuffle of dpc/dataset_3d.py:108-112,
 * bit-exact with Pillow / torchvision arithmetic.  frames [B,F,H,W,3] uint8 (device); out [B,num_seq,3,seq_len,Ho,Wo] float32.
 * tables (device, int32), per clip (Wo + Ho) * (2 + K) + 1 values: xstart[Wo] xcount[Wo] xcoef[Wo][K] ystart[Ho] ycount[Ho]
 *   ycoef[Ho][K] xstep -- separable resampling taps in SOURCE coordinates, 22-bit fixed point (crop, flips, resize folded in);
 * frame_params (device, int32), per frame 10 values: grey channel (-1 none), op[4] (0 brightness 1 contrast 2 saturation 3 hue,
 *   -1 end), factor[4] (float32 bits), hue byte.  mean / stdv: 3 floats each, HOST pointers. */
int dpc_augment_clips(const uint8_t* frames, const int32_t* tables, const int32_t* frame_params, const float* mean,
                      const float* stdv, float* out, int B, int F, int H, int W, int Ho, int Wo, int K, int num_seq,
                      int seq_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPC_B200_H */
