/* rnvp.h -- C ABI of the B200-native RealNVP hot path (librnvp_b200.so).
 *
 * Drop-in boundary for hse-cs/probaforms' RealNVP path.  The reference has no
 * native code: its hot path is Python calling torch eager ops.  Each entry
 * point below replaces the reference interface cited next to it (paths are
 * relative to the reference repo root); INTEGRATION.md shows the ctypes stub a
 * maintainer would add on the reference side.
 *
 * Conventions
 *  - plain C types only; every pointer named d_* is a DEVICE pointer to
 *    contiguous fp32 (or int64 where stated) memory owned by the caller (torch);
 *    the library never allocates or frees caller-visible memory.  The opaque
 *    descriptor owns a few KB of device tables (its per-tile op programs).
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it.  The per-tile programs of the
 *    FP32 kernels for the whole flow are built and uploaded by rnvp_desc_create (synchronous, once), so
 *    rnvp_forward / rnvp_inverse / rnvp_sample / rnvp_backward / rnvp_adam_step over the full layer range
 *    neither allocate nor synchronise and are CUDA-graph capturable; a call on a layer SUB-range
 *    (RealNVPLayer.f / .g) builds its program on first use (one cudaMalloc + synchronous copy, then cached).
 *  - the only process-global mutable state is the development trace pointer of rnvp_debug_set_trace;
 *    everything else lives in the descriptor (guarded by a mutex) or on the caller's stream.
 *  - every function returns 0 on success, a negative RNVP_E* code for argument /
 *    planning errors, or a positive cudaError_t.  rnvp_last_error() returns a
 *    thread-local message.  Nothing throws, exits, or falls back to the CPU.
 *  - parameters cross the boundary in the reference's own layout: one flat fp32
 *    buffer in `nf.parameters()` order, per coupling layer i
 *    nn_t.0.weight [h0, D+Cd], nn_t.0.bias [h0], ..., nn_t.{2n}.weight [D, h_last],
 *    nn_t.{2n}.bias [D], then the same for nn_s  (probaforms/models/realnvp.py:69-70,
 *    gen_network realnvp.py:19-43).  rnvp_pack_params() converts it to the
 *    kernel-private mask-compacted "packed" layout; gradients come back through
 *    rnvp_unpack_grads() with exact zeros on masked rows/columns, as autograd
 *    produces them in the reference.
 */
#ifndef RNVP_H
#define RNVP_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rnvp_desc rnvp_desc;

enum {
  RNVP_OK = 0,
  RNVP_EINVAL = -1,      /* bad argument */
  RNVP_ESHAPE = -2,      /* flow shape cannot be planned (e.g. shared-memory budget) */
  RNVP_ENODEVICE = -3,   /* no sm_100 device / wrong architecture */
  RNVP_EALLOC = -4
};
enum { RNVP_ACT_TANH = 1, RNVP_ACT_RELU = 2 };   /* realnvp.py:32-37 */

/* Build the descriptor of a flow on the CURRENT CUDA device.
 * Replaces RealNVP._model_init's layer construction (realnvp.py:195-204):
 * L coupling layers over var_size=D, cond_size=Cd (0 when C is None), conditioner
 * hidden widths hidden[0..n_hidden-1], mask_i[j] = (j+i)%2 (realnvp.py:199). */
int rnvp_desc_create(int D, int Cd, int L, int n_hidden, const int* hidden, int act, rnvp_desc** out);
void rnvp_desc_destroy(rnvp_desc* d);

/* sizes, in floats */
int64_t rnvp_param_count(const rnvp_desc* d);    /* P: flat reference-layout parameters */
int64_t rnvp_packed_count(const rnvp_desc* d);   /* kernel-private packed layouts (all kernel families) */
int64_t rnvp_grad_count(const rnvp_desc* d);     /* floats of the packed gradient accumulator d_gpacked */
/* bytes of scratch rnvp_backward needs for a batch of N rows: the per-CTA x_T stash of the fused FP32 kernel
 * (independent of N, stays L2 resident); on the tensor-core fit paths the per-layer (x_T, s) stash plus the
 * activation records, Npad*L*(D + rnvp_wgrad_record_floats)*4 bytes (Npad = N rounded up to whole row tiles) */
int64_t rnvp_workspace_bytes(const rnvp_desc* d, int64_t N);
/* offsets[2*k], offsets[2*k+1] = (float offset, numel) of the k-th tensor of nf.parameters();
 * n = 4*(n_hidden+1)*L entries pairs.  Returns the number of tensors. */
int rnvp_param_tensors(const rnvp_desc* d, int64_t* offsets, int max_tensors);
/* plan introspection: mode 0 forward, 1 inverse, 2 fused forward+backward, 3 backward-only sweep (used after the
 * tcgen05 forward), 4 = "is the whole fit step on the tensor cores" (kernel_family 2 iff yes); rows per CTA tile, shared
 * memory, ops per tile of the FP32 program, kernel family (0 tile, 1 small-flow, 2 tcgen05) */
int rnvp_plan_info(const rnvp_desc* d, int mode, int* tile_rows, int* smem_bytes, int* n_ops, int* kernel_family);

/* flat (reference layout) -> packed; run after every parameter update */
int rnvp_pack_params(const rnvp_desc* d, const float* d_flat, float* d_packed, void* stream);
/* packed gradient accumulator -> flat reference layout (masked entries written as exact 0) */
int rnvp_unpack_grads(const rnvp_desc* d, const float* d_gpacked, float* d_gflat, void* stream);

/* Forward pass = body of NormalizingFlow.log_prob before the mean (nflow.py:107-115) over
 * layers [layer_begin, layer_end): z, per-row sum of log|det J|, and
 * logp = logdet + MultivariateNormal(0,I).log_prob(z).  Any of d_z/d_logdet/d_logp may be NULL.
 * layer_end - layer_begin == 1 is RealNVPLayer.f (realnvp.py:73-101).  d_C NULL iff Cd == 0.
 * d_idx (int64[N], may be NULL) gathers rows: row r reads X[idx[r]], C[idx[r]]. */
int rnvp_forward(const rnvp_desc* d, const float* d_packed, const float* d_X, const float* d_C,
                 const int64_t* d_idx, int64_t N, int layer_begin, int layer_end,
                 float* d_z, float* d_logdet, float* d_logp, void* stream);

/* Inverse pass = NormalizingFlow.sample after the prior draw (nflow.py:142-143): layers
 * [layer_begin, layer_end) applied in REVERSE order to the latent rows d_Y (the prior sample);
 * one layer is RealNVPLayer.g (realnvp.py:104-129). */
int rnvp_inverse(const rnvp_desc* d, const float* d_packed, const float* d_Y, const float* d_C,
                 int64_t N, int layer_begin, int layer_end, float* d_X, void* stream);

/* NormalizingFlow.sample in one launch (nflow.py:141-143): the prior draw X = prior.sample((n,)) is generated inside the
 * inverse kernel instead of being written to and read back from device memory.  Latent element (row, j) is a pure
 * function of (seed, row_offset + row, j) -- Philox4x32-10 + Box-Muller, csrc/rnvp_philox.cuh -- so a row block gives
 * the same rows whichever GPU or launch produces it: shard a request of n rows over GPUs by passing each shard its
 * first global row as row_offset (no communication).  rnvp_inverse stays the parity mode (caller-supplied noise). */
int rnvp_sample(const rnvp_desc* d, const float* d_packed, const float* d_C, int64_t N, uint64_t seed,
                int64_t row_offset, float* d_X, void* stream);

/* Fused forward + backward of  out = scale * sum_rows logp(row)  (loss.backward() of
 * loss = -nf.log_prob(X, C), realnvp.py:246-250, is scale = -1/N).  Weight gradients are ACCUMULATED into
 * d_gpacked (caller zeroes it), sum_rows logp is accumulated into d_logp_sum (1 float, may be NULL), per-row
 * logp optionally written.  Kernel families: small flows and the generic FP32 tile kernel recompute the
 * hidden activations in the backward sweep (nothing but x_T per layer is kept); the tensor-core fit paths
 * (D = 32 flows with H <= 128; D = 64 / 128 flows with H a multiple of 128) hand h, u and delta2 of every
 * (layer, row) to the weight-gradient sweep through activation records in d_workspace (see
 * rnvp_wgrad_sweep), which is why rnvp_workspace_bytes grows with N there. */
int rnvp_backward(const rnvp_desc* d, const float* d_packed, const float* d_X, const float* d_C,
                  const int64_t* d_idx, int64_t N, float scale, float* d_gpacked, float* d_logp_sum,
                  float* d_logp, void* d_workspace, int64_t workspace_bytes, void* stream);

/* torch.optim.Adam step (realnvp.py:205-207, 251; betas/eps as constructed there) on the flat
 * parameters, refreshing d_packed in the same pass.  The gradient is read from the packed
 * accumulator d_gpacked, or, when d_gflat_in is non-NULL, from a reference-layout gradient (the
 * .grad tensors autograd filled); grad_scale multiplies it first (e.g. 1/world_size after an
 * all-reduce).  step is the 1-based step count; scalar hyper-parameters are doubles because torch
 * does the bias-correction arithmetic in Python doubles.  d_gflat_out (may be NULL) receives the
 * reference-layout gradient.  zero_gpacked != 0 re-zeroes the accumulator entries it consumed
 * (opt.zero_grad(), realnvp.py:249).  If d_loss_src is non-NULL, *d_loss_dst = *d_loss_src *
 * loss_scale (loss_history entry, realnvp.py:254) and *d_loss_src is re-zeroed with the gradients. */
int rnvp_adam_step(const rnvp_desc* d, float* d_flat, float* d_packed, float* d_gpacked,
                   const float* d_gflat_in, float* d_exp_avg, float* d_exp_avg_sq, float* d_gflat_out,
                   float grad_scale, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int64_t step, int zero_gpacked, float* d_loss_src,
                   float* d_loss_dst, float loss_scale, void* stream);

/* One epoch of the reference's fit loop (realnvp.py:238-254) as ONE call: for every consecutive slice of `batch_size`
 * entries of the epoch's row order d_perm[n] (the last partial batch is kept): rnvp_backward with scale = -1/nb, then
 * rnvp_adam_step (step count step0 + 1, step0 + 2, ...), the step's loss written to d_losses[s].  Same arithmetic as the
 * two calls it wraps, bit for bit; it removes the host-side per-step overhead, which dominates README-sized batches, and
 * for small flows whose batch fits one 32-row tile (the reference's default batch_size) each step is ONE launch: the
 * fit kernel applies the Adam update, refreshes d_packed and hands the loss off itself (the step's gradient is already in
 * its shared memory; 16 us per step on B200 against 23 us for the two launches).  Single GPU (no all-reduce between the
 * two).  d_gpacked and *d_loss_slot must be zero on entry and are zero again on return. */
int rnvp_fit_epoch(const rnvp_desc* d, float* d_flat, float* d_packed, float* d_gpacked, float* d_exp_avg, float* d_exp_avg_sq,
                   const float* d_X, const float* d_C, const int64_t* d_perm, int64_t n, int64_t batch_size, double lr,
                   double beta1, double beta2, double eps, double weight_decay, int64_t step0, float* d_loss_slot,
                   float* d_losses, void* d_workspace, int64_t workspace_bytes, void* stream);

/* Weight-gradient sweep from stored activations (last part of a fit step on the tcgen05 path): for every layer
 * dW1 += delta1^T u, db1 += sum delta1, dW2 += delta2^T h, db2 += sum delta2 (sums over rows), accumulated into
 * d_gpacked, with delta1 = (delta2 W2) * act'(h) recomputed from d_packed.  One record of
 * rec = rnvp_wgrad_record_floats(d) floats per (layer, row):
 * h [2][H] (nn_t | nn_s) | u = [x_K, c, 0..] (ceil8(D/2+Cd)) | delta2 [2][D/2],
 * stored in blocks of 32 rows as d_records[L][Npad/32][rec/4][32][4]: float4 column group q of row r of a block sits
 * in slot (r ^ (q & 7)).  Npad is a multiple of 32; padding rows must hold zeros in delta1 / delta2.
 * Flows whose fit step runs on the tensor cores: D = 32 with H <= 128 (multiple of 16), D = 64 / 128 with H a
 * multiple of 128 (one 128-lane block of hidden units per CTA, rnvp_wgrad_tc.cu). */
int rnvp_wgrad_record_floats(const rnvp_desc* d);
int rnvp_wgrad_sweep(const rnvp_desc* d, const float* d_packed, int64_t Npad, const float* d_records, float* d_gpacked,
                     void* stream);

/* Kernel-family selection for rnvp_forward / rnvp_inverse: 0 = auto (tcgen05 TF32x3 kernels where the shape is
 * eligible -- one hidden layer, D/2 in {16,32} --, else the small-flow or FP32 tile kernels), 1 = FP32-FMA kernels
 * only, 2 = same as auto.  rnvp_plan_info reports the family chosen (0 tile, 1 small-flow, 2 tcgen05). */
int rnvp_set_path(rnvp_desc* d, int path);

/* Host-side row order of one epoch, produced incrementally (replaces the per-epoch DataLoader(shuffle=True) of
 * realnvp.py:237: RandomSampler -> torch.randperm(n, generator=Generator().manual_seed(seed)); bit-identical to it).
 * `out` is a caller-owned HOST buffer of n int64 (pinned memory recommended).  rnvp_perm_advance finalises
 * out[0 .. upto) and returns the number of final entries, so the first batches can be consumed while a helper thread is
 * still shuffling the tail.  n must be below 2^32/20 (torch uses another scheme beyond: RNVP_ESHAPE).  No device work. */
typedef struct rnvp_perm rnvp_perm;
int rnvp_perm_create(uint64_t seed, int64_t n, int64_t* out, rnvp_perm** p);
int64_t rnvp_perm_advance(rnvp_perm* p, int64_t upto);
void rnvp_perm_destroy(rnvp_perm* p);

/* Host-side ingestion / egress helpers (no device work; HOST pointers).  rnvp_host_gather_rows writes
 * dst[r][0..width) = (float) src[idx ? idx[r] : row0 + r][0..width) for r in [0, n): the rows of one optimisation step,
 * converted from the caller's float64 (src_is_f64 != 0) or float32 array -- torch.tensor(X, dtype=float32) of
 * realnvp.py:226-228 fused with the batch gather of realnvp.py:237 -- into a (pinned) staging buffer, on up to `threads`
 * host threads.  The staging buffer is written with non-temporal stores when it is 16-byte aligned and width % 4 == 0:
 * its next reader is the GPU's DMA engine, and an upload whose source lines are dirty in many cores' caches runs at a
 * quarter of the bus rate.  rnvp_host_copy is a multi-threaded memcpy (the .cpu().numpy() of realnvp.py:281 into a
 * fresh pageable array; RealNVP.sample uses it only beyond the pinned result pool's cap). */
int rnvp_host_gather_rows(const void* src, int src_is_f64, int64_t width, const int64_t* idx, int64_t row0, int64_t n,
                          float* dst, int threads);
/* X and C rows of one step in one pass over idx (src_c may be NULL) */
int rnvp_host_gather_xc(const void* src_x, int x_is_f64, int64_t width_x, const void* src_c, int c_is_f64, int64_t width_c,
                           const int64_t* idx, int64_t row0, int64_t n, float* dst_x, float* dst_c, int threads);
int rnvp_host_copy(void* dst, const void* src, int64_t bytes, int threads);

/* Development aid: when d_buf (device, 4*2048*2 int64) is non-null, CTA 0 of the tcgen05 fit kernel logs (tag, clock64)
 * pairs of its backward sweep (tile-0 epilogue, tile-1 epilogue, and their two MMA issuers); tools/trace_mma.py prints the timeline. */
int rnvp_debug_set_trace(void* d_buf);

/* Tensor-core primitive self-test (tcgen05.mma kind::tf32, A in TMEM, B in shared memory):
 * D[128,N] = A[128,K] * B[N,K]^T on device buffers; passes = 1 (plain TF32) or 3 (split, fp32-grade). */
int rnvp_mma_selftest(const float* d_A, const float* d_B, float* d_D, int N, int K, int passes, void* stream);

const char* rnvp_last_error(void);
int rnvp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RNVP_H */
