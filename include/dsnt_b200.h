/*
 * dsnt_b200.h -- C ABI of libdsnt_b200.so, the sm_100a implementation of the DSNT head hot path.
 *
 * This is the drop-in boundary.  The reference (anibali/dsnt-pose2d) is pure
 * Python: its "FFI" for this path is the set of tensor functions in src/dsnt/nn.py plus the two
 * model shims in src/dsnt/model.py.  Each entry point below names the reference lines it replaces.
 * The Python host side (dsnt_pose2d_b200/nn.py, head.py) binds these with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer on the current device
 *   - the library allocates nothing and keeps no state besides a thread-local error string
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*); no host sync,
 *     so a sequence of calls is CUDA-graph capturable
 *   - return value 0 = enqueued; negative = rejected (see dsnt_b200_last_error())
 *   - heatmaps are N contiguous images of H rows x W columns (NCHW conv output viewed as [N=B*C,H,W]);
 *     element (n,i,j) lives at base + (n*H*W + i*W + j)
 *   - pixel-centre coordinates x_j = (2j+1)/W - 1, y_i = (2i+1)/H - 1   (src/dsnt/nn.py:30-37)
 *   - all per-heatmap scalars (coords, stats, terms, targets, masks, gradients) are float32
 */
#ifndef DSNT_B200_H_
#define DSNT_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DSNT_B200_VERSION 100

#if defined(__GNUC__)
#define DSNT_API __attribute__((visibility("default")))
#else
#define DSNT_API
#endif

/* element type of the heatmap tensors (math is always fp32) */
#define DSNT_DTYPE_F32 0
#define DSNT_DTYPE_BF16 1

/* regulariser selector: src/dsnt/model.py:52-61 ('none'|'var'|'kl'|'js'|'mse') */
#define DSNT_REG_NONE 0
#define DSNT_REG_VAR 1
#define DSNT_REG_KL 2
#define DSNT_REG_JS 3
#define DSNT_REG_MSE 4

/* heatmap pre-activation: src/dsnt/model.py:24-45 ('softmax'|'thresholded_softmax'|'abs'|'relu'|'sigmoid') */
#define DSNT_PREACT_SOFTMAX 0
#define DSNT_PREACT_TSOFTMAX 1
#define DSNT_PREACT_ABS 2
#define DSNT_PREACT_RELU 3
#define DSNT_PREACT_SIGMOID 4

/* per-heatmap statistics saved by the forward for the reduction-free backward */
#define DSNT_MAX_STACKS 16 /* stacks (hourglass outputs) one *_stacked call can cover */
#define DSNT_MAX_RANKS 16         /* ranks of one node that exchange partial sums through peer memory */

#define DSNT_STATS_K 8
/*   logits input : [0] m*log2(e)  [1] 1/S   [2] mu_x [3] mu_y [4] v_x [5] v_y [6] c_reg = sum P r [7] Gaussian normaliser 1/(sum G^ + 1e-24)
 *   heatmap input: [0] sum P      [1] unused [2..7] as above                                                        */

/* flags */
#define DSNT_FLAG_STRICT_NAN 1 /* reproduce the reference's NaN gradient when coords == target (sqrt'(0)); default: 0-gradient */
#define DSNT_FLAG_NO_EUCLID 2  /* dsnt_head_bwd: g_loss feeds the regulariser only (standalone *_reg_loss) */

/* error codes */
#define DSNT_OK 0
#define DSNT_ERR_BAD_ARG (-1)
#define DSNT_ERR_UNSUPPORTED (-2)
#define DSNT_ERR_LAUNCH (-3)

DSNT_API int dsnt_b200_version(void);
DSNT_API const char* dsnt_b200_last_error(void);

/*
 * Fused forward of the head for N heatmaps.
 *   replaces: F.softmax over H*W (src/dsnt/model.py:24-30,44-45) when input_is_logits != 0,
 *             dsnt / generate_xy / expectation_2d (src/dsnt/nn.py:25-78),
 *             the per-heatmap part of euclidean_loss (src/dsnt/nn.py:112-114),
 *             make_gauss + _kl_2d/_js_2d/MSE/variance terms (src/dsnt/nn.py:168-216,232-233,250-251,268-270,288-296).
 *   z         [N,H,W] logits, or normalised/un-normalised heatmaps P when input_is_logits == 0
 *   target    [N,2] (x,y) in normalised units, or NULL (then dist = 0 and reg must be NONE or VAR)
 *   sigma     target std-dev in NORMALISED units (= 2*hm_sigma/W, src/dsnt/model.py:49)
 *   coords    [N,2] out: (E[x], E[y])
 *   stats     [N,DSNT_STATS_K] out, or NULL when no backward will follow
 *   terms     [N,2] out: (Euclidean distance to target, regulariser value D), or NULL
 *   variant   0 = automatic kernel choice; >0 forces an implementation variant (benchmarks/tests)
 */
DSNT_API int dsnt_head_fwd(const void* z, int dtype, int input_is_logits, long n, int H, int W,
                  const float* target, int reg, float sigma,
                  float* coords, float* stats, float* terms, int variant, void* stream);

/*
 * Fused, reduction-free backward: recomputes softmax from the saved statistics and writes
 * dL/dz directly.   replaces: the autograd replay of everything above (src/dsnt/nn.py:25-298 backward).
 *
 * Per heatmap n the upstream gradients are assembled on the device from up to three sources
 *   a_n   = g_coords[n,0] + g_loss * w_n * (mu_x - t_x)/d_n       (second term iff g_loss != NULL)
 *   b_n   = g_coords[n,1] + g_loss * w_n * (mu_y - t_y)/d_n
 *   rho_n = g_reg[n]      + g_loss * w_n * reg_coeff
 *   w_n   = (mask ? mask[n] : 1) / denom[0]                       (masked_average, src/dsnt/nn.py:81-94)
 * and dz = P (a x + b y + rho r - c) for logits, dz = a x + b y + rho r for heatmap input.
 *   g_coords [N,2] or NULL; g_reg [N] or NULL; g_loss, denom: device scalars or NULL (both or neither)
 *   dz        [N,H,W] out, same dtype as z
 */
DSNT_API int dsnt_head_bwd(const void* z, int dtype, int input_is_logits, long n, int H, int W,
                  const float* target, const float* mask, const float* stats,
                  const float* g_coords, const float* g_reg, const float* g_loss, const float* denom,
                  float reg_coeff, int reg, float sigma, int flags,
                  void* dz, int variant, void* stream);

/*
 * The whole training step of the head in ONE pass over the logits (csrc/head_step.cuh): every heatmap is brought
 * into shared memory by a TMA bulk copy, reduced (forward), overwritten with dL/dz (backward) and bulk-stored --
 * 2*H*W*sizeof bytes per heatmap instead of the 3*H*W*sizeof of dsnt_head_fwd + dsnt_head_bwd.
 *   replaces: forward_part2 + forward_loss + loss.backward() of one step (src/dsnt/bin/train.py:355-381) for the
 *             softmax pre-activation; the backward's only cross-heatmap input, the denominator of masked_average
 *             (src/dsnt/nn.py:88-92), depends on the mask alone and is computed first by dsnt_mask_count.
 *   denom     DEVICE scalar max(sum mask, 1) (out[3] of dsnt_mask_count; all-reduce out[2] + dsnt_combine_loss first
 *             when the batch is sharded);  g_loss: DEVICE scalar d(loss) or NULL = 1
 *   dz        [N,H,W] out = g_loss * d(euclid + reg_coeff*reg)/dz, same dtype as z;  coords/stats/terms as dsnt_head_fwd
 *             (feed terms to dsnt_finish_loss for the loss value; stats allow a later dsnt_head_bwd with other gradients)
 *   returns DSNT_ERR_UNSUPPORTED when fewer than 4 heatmaps fit in shared memory or the layout has no 16-byte vectors
 *   (dsnt_head_step_supported tells in advance); the caller then uses the two-kernel path.
 * dsnt_mask_count: out[2] = sum mask (n when mask is NULL), out[3] = max(out[2], 1); workspace as dsnt_finish_loss.
 */
DSNT_API int dsnt_head_step_supported(int dtype, int H, int W);
/* ... for a given regulariser: the shared-memory ring above, or the cluster kernel below (256x256). */
DSNT_API int dsnt_head_step_supported_reg(int dtype, int H, int W, int reg);
/* 256x256 (BASELINE config 5) with no, the variance, the JS or the MSE regulariser is served by a cluster of two CTAs
 * (four where two 1024-thread CTAs cannot be co-scheduled), each holding a part of the heatmap in registers, partial results
 * exchanged through distributed shared memory, most of the gradient written back by bulk stores (csrc/step_pair.cu); KL
 * at this size keeps the two-kernel path.  fp32 runs at the HBM roofline; bf16 is served too but bound by the SM and slower
 * than dsnt_head_fwd + dsnt_head_bwd, so the Python dispatcher does not take it by default. */
DSNT_API int dsnt_head_step_pair_supported(int dtype, int H, int W, int reg);
DSNT_API int dsnt_head_step(const void* z, int dtype, long n, int H, int W, const float* target, const float* mask,
                            const float* denom, const float* g_loss, float reg_coeff, int reg, float sigma, int flags,
                            float* coords, float* stats, float* terms, void* dz, void* stream);
DSNT_API int dsnt_mask_count(const float* mask, long n, float* out, float* workspace, void* stream);
/*
 * The same step in ONE launch (csrc/head_step2.cuh): every CTA adds up its slice of the mask while its first bulk loads
 * are in flight and draws a ticket; the CTA with the last ticket publishes the count, which every warp picks up only
 * before its first BACKWARD (the forward does not need the denominator) -- no dsnt_mask_count, no grid barrier; and the
 * CTA that finishes last composes the loss (no dsnt_finish_loss): out[0..7] exactly as dsnt_finish_loss documents,
 * workspace likewise.  The grid (one CTA per SM) is launched with the cooperative attribute, so its co-residency is
 * guaranteed by the driver.  Served for the shapes / regularisers of the shape-specialised kernel
 * (dsnt_head_step_fused_supported: 64x64; JS / MSE: Gaussian window within its register slots, i.e. sigma up to ~1.4 px;
 * KL walks its window and takes any sigma); otherwise
 * DSNT_ERR_UNSUPPORTED and the caller uses the three-launch form.  A sharded batch: dsnt_head_step_fused_peer below.
 * dsnt_finish_trace_offset_bytes: byte offset inside the workspace of eight uint64 %globaltimer stamps the last
 *   single-launch step left there (diagnostics; bench.py's per-rank timeline): [0] kernel start, [1] local mask count known,
 *   [2] count published (sharded: after the exchange with the other ranks), [3] last CTA done, [4] loss block written,
 *   [5] entry of CTA 0 (before its barriers are initialised and its first loads issued).
 */
DSNT_API int dsnt_head_step_fused_supported(int dtype, int H, int W, int reg, float sigma);
DSNT_API int dsnt_finish_trace_offset_bytes(void);
DSNT_API int dsnt_head_step_fused(const void* z, int dtype, long n, int H, int W, const float* target, const float* mask,
                                  const float* g_loss, float reg_coeff, int reg, float sigma, int flags, float* coords,
                                  float* stats, void* dz, float* out, float* workspace, void* stream);
/* x[0..numel) *= *g, skipped entirely (no traffic) when *g == 1: applies an upstream d(loss) != 1 to the gradient
 * dsnt_head_step already wrote, without a host read of *g. */
DSNT_API int dsnt_scale_unless_one(void* x, int dtype, long numel, const float* g, void* stream);

/*
 * Stacked-hourglass forms: ONE launch for all stacks.
 *   replaces: the per-stack Python loops of HourglassHumanPoseModel.forward_part2 / forward_loss
 *             (src/dsnt/model.py:238-246,286-292) over the list hourglass.py:166-177 returns.
 *   z / dz    HOST arrays of n_stacks DEVICE pointers; every stack is [n_per_stack,H,W] of the same dtype
 *   target, mask  [n_per_stack,...] shared by the stacks;  coords/stats/terms/g_coords/g_reg  [n_stacks*n_per_stack,...]
 *   dsnt_finish_loss_stacked: out[6] = sum over stacks of (euclid_s + reg_coeff*reg_s); out[2], out[3] = mask count /
 *   denominator of ONE stack (every stack shares it).  n_stacks <= DSNT_MAX_STACKS.
 */
DSNT_API int dsnt_head_fwd_stacked(const void* const* z, int n_stacks, int dtype, int input_is_logits, long n_per_stack,
                                   int H, int W, const float* target, int reg, float sigma,
                                   float* coords, float* stats, float* terms, int variant, void* stream);
DSNT_API int dsnt_head_bwd_stacked(const void* const* z, void* const* dz, int n_stacks, int dtype, int input_is_logits,
                                   long n_per_stack, int H, int W, const float* target, const float* mask,
                                   const float* stats, const float* g_coords, const float* g_reg, const float* g_loss,
                                   const float* denom, float reg_coeff, int reg, float sigma, int flags, int variant,
                                   void* stream);
/* The single-launch step (dsnt_head_step_fused) over all stacks: dz[s] receives d(sum_s euclid_s + reg_coeff*reg_s)/dz_s,
 * out[] as dsnt_finish_loss_stacked.  Same support rule as dsnt_head_step_fused_supported. */
DSNT_API int dsnt_head_step_fused_stacked(const void* const* z, void* const* dz, int n_stacks, int dtype, long n_per_stack,
                                          int H, int W, const float* target, const float* mask, const float* g_loss,
                                          float reg_coeff, int reg, float sigma, int flags, float* coords, float* stats,
                                          float* out, float* workspace, void* stream);
DSNT_API int dsnt_finish_loss_stacked(const float* terms, const float* mask, long n_per_stack, int n_stacks,
                                      float reg_coeff, float* out, float* workspace, void* stream);

/*
 * The fused head for the other pre-activations of the reference.
 *   replaces: HumanPoseModel._hm_preact (src/dsnt/model.py:24-45) -- thresholded_softmax(z, -0.5)
 *             (src/dsnt/nn.py:119-157), |z|, relu(z), sigmoid(z), each divided by (sum + 1e-12) -- followed by dsnt,
 *             the per-heatmap Euclidean distance and the regulariser exactly as dsnt_head_fwd/bwd, plus their autograd
 *             replay (thresholded softmax: the custom backward out*(g - sum g out), src/dsnt/nn.py:131-139).
 *   P = f(z) / (sum f(z) + eps) is never materialised; P may be exactly 0 and need not sum to 1, so every epsilon
 *   of the reference (1e-24 inside the logs, eps in the normaliser) is kept.
 *   preact     DSNT_PREACT_*; DSNT_PREACT_SOFTMAX with eps = 0 is plain softmax with the epsilons kept
 *   threshold  DSNT_PREACT_TSOFTMAX only: entries below it get probability 0 (the max is over ALL entries)
 *   eps        added to the normaliser sum (the reference uses 1e-12)
 *   stats      [N,DSNT_STATS_K]: [0] max*log2(e) [1] 1/(sum+eps) [2..6] as dsnt_head_fwd
 *              [7] Gaussian normaliser (kl/js/mse) or 1 - sum P (var)
 *   variant    0 = automatic: the tuned streaming kernels (the ones dsnt_head_fwd/bwd use, instantiated per
 *              pre-activation) when the layout qualifies, else the generic epsilon-exact kernels; 1 forces the generic ones
 * Other arguments as dsnt_head_fwd / dsnt_head_bwd.
 */
DSNT_API int dsnt_head_preact_fwd(const void* z, int dtype, int preact, float threshold, float eps, long n, int H, int W,
                                  const float* target, int reg, float sigma, float* coords, float* stats, float* terms,
                                  int variant, void* stream);
DSNT_API int dsnt_head_preact_bwd(const void* z, int dtype, int preact, float threshold, long n, int H, int W,
                                  const float* target, const float* mask, const float* stats, const float* g_coords,
                                  const float* g_reg, const float* g_loss, const float* denom, float reg_coeff, int reg,
                                  float sigma, int flags, void* dz, int variant, void* stream);

/*
 * Inference with flip test-time augmentation, fused in front of the forward-only head.
 *   replaces: src/dsnt/inference.py:36-48 -- hm1, hm2 = heatmaps of [images, mirrored images];
 *             hm2 = reverse_tensor(hm2, -1).index_select(-3, HFLIP_INDICES) (src/dsnt/util.py:201-210,
 *             src/dsnt/data.py:97); hm = (hm1 + hm2)/2; coords = forward_part2(hm) (src/dsnt/model.py:176-183)
 *   z          [2*batch, C, H, W] raw heatmaps: the first `batch` samples are the original images, the last `batch`
 *              their mirrored versions (the torch.cat order of inference.py:36)
 *   flip_perm  [C] int32 DEVICE array: joint c of the mirrored image is joint flip_perm[c] (HFLIP_INDICES); NULL = identity
 *   preact/threshold/eps  as dsnt_head_preact_fwd (softmax: eps 0)
 *   coords     [batch*C, 2] out
 *   avg_out    optional [batch, C, H, W] out (same dtype as z): the averaged raw heatmaps `hm`; NULL = not materialised
 * One launch, both heatmap sets read once (2*H*W*sizeof bytes per output heatmap), no statistics saved.
 * A forward-only head without the flip is dsnt_head_fwd / dsnt_head_preact_fwd with stats = terms = NULL.
 */
DSNT_API int dsnt_flip_tta_fwd(const void* z, int dtype, long batch, int C, int H, int W, const int* flip_perm, int preact,
                               float threshold, float eps, float* coords, void* avg_out, void* stream);

/*
 * Deterministic finishing reduction over the per-heatmap terms (no float atomics, one launch).
 *   replaces: masked_average (src/dsnt/nn.py:81-94) and loss = euclid + reg_coeff*reg (src/dsnt/model.py:145).
 *   terms [N,2] from dsnt_head_fwd; mask [N] or NULL (then every weight is 1 and count = N)
 *   out[0] = sum mask*dist   out[1] = sum mask*D   out[2] = sum mask (count)
 *   out[3] = max(count,1)    out[4] = out[0]/out[3]   out[5] = out[1]/out[3]
 *   out[6] = out[4] + reg_coeff*out[5] (the loss)     out[7] = 0
 *   workspace: dsnt_finish_workspace_bytes() bytes of device scratch, 16-byte aligned, ZEROED ONCE by the
 *   caller before first use; the kernel leaves it ready for the next call (stream-ordered reuse only).
 */
DSNT_API int dsnt_finish_workspace_bytes(void);
DSNT_API int dsnt_finish_loss(const float* terms, const float* mask, long n, float reg_coeff, float* out,
                     float* workspace, void* stream);

/*
 * Sharded batch (one process per GPU of a node): the same reductions with the exchange of the partial sums between
 * the ranks fused into the kernel -- no NCCL launch, no separate combine.  The last CTA stores (out[0], out[1], out[2])
 * into every rank's exchange buffer over NVLink peer mappings (each value as one 64-bit word {float, epoch}: no fence, no
 * flag), waits for all ranks' words in its own buffer and adds them in rank order, so every rank finishes with the same
 * out[0..7] as one process on the whole batch would.
 *   replaces: nothing in the reference (single GPU, src/dsnt/bin/train.py:220); it is what makes masked_average
 *             (src/dsnt/nn.py:81-94) mean "over the global batch" when the batch dimension is sharded.
 *   peers     HOST array of `world` DEVICE pointers; peers[r] is rank r's exchange buffer as mapped into this process
 *             (dsnt_peer_exchange_bytes() bytes each, symmetric memory), zeroed once before first use
 *   epoch     local DEVICE unsigned counter, zeroed once;  error: local DEVICE int, set if a peer never arrived (20 s)
 *   All ranks must issue the same sequence of *_peer calls on the same group (as with any collective).
 */
DSNT_API int dsnt_peer_exchange_bytes(void);
DSNT_API int dsnt_finish_loss_peer(const float* terms, const float* mask, long n_per_stack, int n_stacks, float reg_coeff,
                                   float* out, float* workspace, const void* const* peers, int rank, int world,
                                   unsigned* epoch, int* error, void* stream);
DSNT_API int dsnt_mask_count_peer(const float* mask, long n, float* out, float* workspace, const void* const* peers,
                                  int rank, int world, unsigned* epoch, int* error, void* stream);
/* The single-launch step (dsnt_head_step_fused) on a sharded batch: the CTA that publishes the mask count first exchanges
 * it with the other ranks (hidden behind the first loads and the first forward of every warp, as the count itself), and
 * the CTA with the last ticket exchanges the loss sums at the end -- one launch per rank and step, no collective call.
 * n > 0 on this rank; a rank with an empty shard issues dsnt_mask_count_peer + dsnt_finish_loss_peer instead and meets
 * the others in the same two exchanges. */
DSNT_API int dsnt_head_step_fused_peer(const void* z, int dtype, long n, int H, int W, const float* target, const float* mask,
                                       const float* g_loss, float reg_coeff, int reg, float sigma, int flags, float* coords,
                                       float* stats, void* dz, float* out, float* workspace, const void* const* peers,
                                       int rank, int world, unsigned* epoch, int* error, void* stream);

/* Recompute out[3..6] from (possibly all-reduced) out[0..2]; used after the NCCL all-reduce of the sums. */
DSNT_API int dsnt_combine_loss(float* out, float reg_coeff, void* stream);

/*
 * Per-point Euclidean distance for the standalone euclidean_loss (src/dsnt/nn.py:97-116):
 *   terms[n] = (sqrt(sum_k (actual[n,k]-target[n,k])^2), 0), k < d  -- the [N,2] layout dsnt_finish_loss averages.
 * Backward:  g_actual[n,k] = g_loss * w_n * (actual-target)/dist,  w_n = (mask ? mask[n] : 1)/denom
 *   (0 instead of NaN at dist == 0 unless DSNT_FLAG_STRICT_NAN)
 */
DSNT_API int dsnt_euclid_fwd(const float* actual, const float* target, long n, int d, float* terms, void* stream);
DSNT_API int dsnt_euclid_bwd(const float* actual, const float* target, const float* terms, const float* mask,
                             const float* g_loss, const float* denom, long n, int d, int flags, float* g_actual,
                             void* stream);

/*
 * (Thresholded) softmax over the last dimension of a [rows, len] matrix.
 *   replaces: ThresholdedSoftmax.forward/backward (src/dsnt/nn.py:119-139) and, with threshold = -inf and
 *   eps = 0, softmax_2d / flat_softmax (src/dsnt/nn.py:160-165).  The max is taken over ALL entries.
 *   out = exp(x - max) * (x >= threshold) / (sum + eps);   dx = out * (g - sum(g*out))
 */
DSNT_API int dsnt_tsoftmax_fwd(const void* x, int dtype, long rows, long len, float threshold, float eps,
                      void* out, void* stream);
DSNT_API int dsnt_tsoftmax_bwd(const void* out, const void* g, int dtype, long rows, long len, void* dx, void* stream);

/*
 * Normalised separable Gaussians (make_gauss, src/dsnt/nn.py:168-205; note width before height).
 *   mu [N,2]; out [N,H,W] float32.   Backward wrt mu: dmu [N,2] from g [N,H,W].
 */
DSNT_API int dsnt_make_gauss_fwd(const float* mu, long n, int W, int H, float sigma, float* out, void* stream);
DSNT_API int dsnt_make_gauss_bwd(const float* mu, const float* g, long n, int W, int H, float sigma, float* dmu,
                        void* stream);
/*
 * d(loss)/d(mu_t) of kl / js / mse_reg_loss: in the reference the target Gaussian is built by make_gauss inside the
 * regulariser (src/dsnt/nn.py:232,250,268) and is differentiable w.r.t. its centres (:170-205), so targets that require
 * grad receive one.  dmu[n] = *g_loss * reg_coeff * (mask ? mask[n] : 1) / *denom * dD_n/dmu_n.
 *   z: normalised heatmaps P (input_is_logits = 0) or logits with the forward's stats [N,DSNT_STATS_K] (= 1)
 *   reg: DSNT_REG_KL / JS / MSE (the variance regulariser does not depend on mu_t: src/dsnt/nn.py:274-298)
 */
DSNT_API int dsnt_reg_dmu(const void* z, int dtype, int input_is_logits, long n, int H, int W, const float* stats,
                          const float* mu, const float* mask, const float* g_loss, const float* denom, float reg_coeff,
                          int reg, float sigma, float* dmu, void* stream);


/*
 * The 'gauss' output strategy helpers, float32 arithmetic identical to the reference's.
 *
 * dsnt_draw_gaussians   replaces: encode_heatmaps (src/dsnt/util.py:129-148) and draw_gaussian (:70-126)
 *   centres   [N,2] float32 DEVICE: normalised (x, y) coords (centres_are_pixels == 0: converted and rounded to the
 *             nearest pixel exactly as util.py:133-144) or pixel coordinates (centres_are_pixels != 0: int() truncation)
 *   sigma     std-dev in pixels;  clip_size: side of the square draw region (encode_heatmaps uses 7), <= 0 = unclipped
 *   normalize divide the drawn region by its sum (util.py:122-125)
 *   out       [N,H,W] float32: the Gaussian inside the draw window, 0 elsewhere (encode_heatmaps starts from zeros)
 *
 * dsnt_decode_heatmaps  replaces: get_preds + decode_heatmaps (src/dsnt/util.py:151-198)
 *   hm [N,H,W] (fp32 or bf16) -> coords [N,2] float32 normalised; arg-max with the first maximum winning, (0,0) pixel
 *   when the maximum is <= 0, y = idx / H as in util.py:163, optional quarter-pixel offset toward the higher neighbour.
 */
DSNT_API int dsnt_draw_gaussians(const float* centres, int centres_are_pixels, long n, int W, int H, double sigma,
                                 double clip_size, int normalize, float* out, void* stream);
DSNT_API int dsnt_decode_heatmaps(const void* hm, int dtype, long n, int H, int W, int use_neighbours, float* coords,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSNT_B200_H_ */
