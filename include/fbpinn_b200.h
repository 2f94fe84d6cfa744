/*
 * fbpinn_b200.h — C ABI of libfbpinn_b200.so: the FBPINN per-step subdomain evaluation on B200 (sm_100a).
 *
 * This is the drop-in boundary for the ONE hot path of benmoseley/FBPINNs (v0.2.0, JAX): the work that
 * the reference performs inside the jitted `FBPINN_update` (fbpinns/trainers.py:285-296) and in the index
 * construction it depends on.  The reference has no native/FFI layer of its own; the seam these entry
 * points replace is
 *      FBPINN_forward(all_params, x_batch, takes, model_fns, jmaps) -> ujs     fbpinns/trainers.py:197-203
 * under value_and_grad (:292) plus the optimiser lines (:294-295), and, at active-set changes,
 *      decomposition.inside_points / inside_models                              fbpinns/decompositions.py:201-215
 *      get_inputs / _get_update_inputs index algebra                            fbpinns/trainers.py:332-391, 544-571
 * INTEGRATION.md shows the reference-side binding (jax.ffi custom call + custom_vjp, or ctypes).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  Every pointer named d_* is a DEVICE pointer owned by the
 *    caller; the library never frees or retains caller memory beyond the call (handles excepted).
 *  - Every function returns 0 on success, non-zero on failure; fbp_last_error() gives the message of the
 *    last failure on the calling thread.  Nothing throws across the boundary.  There is no CPU fallback.
 *  - `stream` is a cudaStream_t passed as void*.  All work is enqueued on it; step-path functions
 *    (window_sums, forward, reduce_*, backward, adam_step, gather/scatter) never allocate, never
 *    synchronise and are CUDA-graph capturable.  Index-construction functions (fbp_takes_*) may
 *    allocate scratch and synchronise the stream; they run only when the active set changes.
 *  - float32 values, int32 indices (JAX defaults, x64 disabled in the reference).
 *
 * Data layout (all row-major, contiguous)
 *  - packed parameters  d_params [m][P]:  per subdomain, for each linear layer l: W_l (out x in, the
 *    reference's (out,in) leaf, fbpinns/networks.py:55) then b_l (out).  P = sum_l out_l*(in_l+1).
 *  - static subdomain record d_sub_static [m][2*xd+3]: xmin[xd], xmax[xd], window flag, unnorm mu, unnorm sd
 *    (float32 casts of fbpinns/decompositions.py:135-181 `params[0,1,4,5]`).
 *  - jet components: component 0 is the value; a component of order 1 is d/dx_k; order 2 is d2/dx_k dx_l.
 *    The component set must be closed (an order-2 component needs both its order-1 components) and comes
 *    from the `jmaps` trie of the constraint's required_ujs (fbpinns/trainers.py:63-107).
 *  - per-point jets d_ujets [n][C*ud], index c*ud+o.   per-pair numerator jets d_pair_out [s][C*ud] in
 *    SUBDOMAIN-SORTED pair order.
 *  - takes (see fbp_takes_emit): reference-order arrays (m_take, n_take, p_take, np_take: bit-exact with
 *    fbpinns/trainers.py:336-391) plus the subdomain-sorted view the kernels consume.
 */
#ifndef FBPINN_B200_H
#define FBPINN_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FBP_MAX_LAYERS 16   /* linear layers */
#define FBP_MAX_XD 3
#define FBP_MAX_UD 4
#define FBP_MAX_COMP 10     /* 1 + xd + xd*(xd+1)/2 at xd = 3 */

#define FBP_BWD_ACCUMULATE 1
#define FBP_BWD_NO_REDUCE 2
#define FBP_BWD_REDUCE_ONLY 4
#define FBP_BWD_DIRECT 8

#define FBP_ACT_TANH 0            /* FCN, fbpinns/networks.py:61-68 */
/* The reference's other Network plug-ins (generic kernel family only).  Packed parameter row: per layer W, b, then
 * the layer's per-unit activation parameters (every layer carries them, the last layer's are unused like in the
 * reference's pytree): */
#define FBP_ACT_ADAPTIVE_TANH 1   /* AdaptiveFCN  a * tanh(x / a),   extra: a          networks.py:70-101 */
#define FBP_ACT_SIN 2             /* SIREN        sin(x)                               networks.py:103-133 */
#define FBP_ACT_ADAPTIVE_SIN 3    /* AdaptiveSIREN c * sin(o * x),   extras: c, o      networks.py:135-166 */
#define FBP_ACT_FOURIER_TANH 4    /* FourierFCN: layer 0 is the STATIC feature layer sin(W0 z + b0) with W0 = [omega; omega],
                                   * b0 = [0; pi/2] (= [sin, cos] features, no gradient), then tanh layers  networks.py:168-194 */
#define FBP_WINDOW_COSINE 0 /* windows.cosine, fbpinns/windows.py:25-35 */

typedef struct fbp_plan fbp_plan;                    /* opaque */
typedef struct fbp_takes_builder fbp_takes_builder;  /* opaque */

typedef struct fbp_plan_desc {
    int32_t xd;                                /* input dimension, 1..FBP_MAX_XD */
    int32_t ud;                                /* output dimension, 1..FBP_MAX_UD */
    int32_t n_layers;                          /* number of linear layers = len(layer_sizes)-1 */
    int32_t layer_sizes[FBP_MAX_LAYERS + 1];   /* FCN layer_sizes, fbpinns/networks.py:43 */
    int32_t activation;                        /* FBP_ACT_* */
    int32_t window;                            /* FBP_WINDOW_* */
    int32_t n_comp;                            /* C, jet components incl. the value */
    int32_t comp_k[FBP_MAX_COMP];              /* first axis, -1 for the value */
    int32_t comp_l[FBP_MAX_COMP];              /* second axis, -1 unless order 2 */
} fbp_plan_desc;

/* Kernel-facing view of one constraint's takes (device pointers, see fbp_takes_emit). */
typedef struct fbp_takes_view {
    int64_t n;                   /* points in this constraint's x_batch */
    int64_t s;                   /* (point, subdomain) pairs */
    int64_t q;                   /* unique (point, pou) rows = len(np_take) */
    int64_t s_active;            /* pairs of the m_active leading subdomains = sub_off[m_active] */
    int32_t m_all;               /* subdomains taking part = len(all_ims) (active first, then fixed) */
    int32_t m_active;            /* leading subdomains that are trained */
    int32_t npou;                /* global number of partitions of unity */
    const int32_t* d_m_take;     /* [s] reference order: subdomain position of each pair (trainers.py:372) */
    const int32_t* d_np_take;    /* [q] reference order: point of each row (trainers.py:385) */
    const int32_t* d_sub_ids;    /* [m_all] position -> global subdomain index (all_ims) */
    const int32_t* d_sub_off;    /* [m_all+1] subdomain-sorted pair ranges */
    const int32_t* d_spair_point;/* [s] point index of each subdomain-sorted pair */
    const int32_t* d_spair_row;  /* [s] row (p_take value) of each subdomain-sorted pair */
    const int32_t* d_spair_sub;  /* [s] subdomain position of each subdomain-sorted pair */
    const int32_t* d_pos;        /* [s] reference-order pair index -> subdomain-sorted pair index */
    const int32_t* d_row_off;    /* [q+1] reference-order pair range of each row */
    const int32_t* d_pt_row_off; /* [n+1] row range of each point */
    const int32_t* d_items;      /* [n_items][4] work list (sub position, first pair, pair count, split id) */
    const int32_t* d_sub_item_off;/* [m_all+1] item range of each subdomain position */
    const int32_t* d_item_order_fwd; /* [n_items] launch order of the items (longest first), NULL = identity */
    const int32_t* d_item_order_bwd; /* [n_items_active] launch order of the active items, NULL = identity */
    int32_t n_items;             /* built by the host from d_sub_off (fbp_plan_tile_points) */
    int32_t n_items_active;      /* leading items that belong to active subdomains */
    /* optional (NULL = resolved on the device through d_item_order_*, d_items, d_sub_ids: two more dependent loads per
       work item): the same information flattened into one 16-byte record per launched block, in launch order:
       (first pair, pair count, global subdomain index, item) */
    const int32_t* d_launch_fwd; /* [n_items][4] */
    const int32_t* d_launch_bwd; /* [n_items_active][4] */
} fbp_takes_view;

/* ---- errors / info ---------------------------------------------------------------------------- */
const char* fbp_last_error(void);
int fbp_version(void);
/* Number of kernels this library has launched (or recorded into a CUDA graph being captured) so far. */
int64_t fbp_launch_count(void);
/* Device properties the host uses for grid sizing. Fails if no CUDA device / not sm_100. */
int fbp_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* smem_per_block_optin);

/* ---- plan: network shape + jet spec (replaces the static args model_fns / jmaps of
 *      FBPINN_forward, fbpinns/trainers.py:197, 285) ------------------------------------------- */
int fbp_plan_create(fbp_plan** plan, const fbp_plan_desc* desc);
int fbp_plan_destroy(fbp_plan* plan);
int64_t fbp_plan_param_count(const fbp_plan* plan);       /* P */
int32_t fbp_plan_is_fast(const fbp_plan* plan);           /* 1 if a tiled kernel instance covers this plan */
int32_t fbp_plan_tile_points(const fbp_plan* plan);       /* points per CTA tile (work-list granularity) */
/* Select kernel family: 0 = auto (tiled when available; the tensor forward where its instance has been validated on
 * hardware, unless the environment says FBP_TC_AUTO=0; FBP_TC_AUTO=full adds the tensor reverse kernel), 1 = force generic, 2 = force tiled (error if none),
 * 3 = tensor: the forward hidden-layer GEMMs run on the tcgen05 tensor cores in 3xTF32 (FP32-equivalent accuracy);
 *     needs H = 32, two hidden layers and at most 5 jet components (error otherwise).  The reverse kernel stays tiled.
 * 4 = tensor forward and the warp-specialised tensor reverse kernel (no activation cache; bring-up state). */
int fbp_plan_set_kernel(fbp_plan* plan, int32_t mode);
int32_t fbp_plan_has_tensor(const fbp_plan* plan);        /* 1 if mode 3 is available for this plan */
int32_t fbp_plan_forward_family(const fbp_plan* plan);    /* family fbp_forward will use: 0 generic, 1 tiled, 2 tensor */
int32_t fbp_plan_reverse_family(const fbp_plan* plan);    /* family fbp_backward will use: 0 generic, 1 tiled, 2 tensor */
/* Scratch floats the generic kernels need per pair (0 for tiled plans in auto mode). */
int64_t fbp_plan_scratch_per_pair(const fbp_plan* plan);
/* Floats per pair of the optional activation cache (0 if the plan's kernels do not use one).  When a cache of
 * s * that many floats is passed to fbp_forward, the tiled forward kernel saves the jets of the last hidden layer
 * in it and fbp_backward (given the same buffer, same parameters) loads them back with TMA bulk copies instead of
 * recomputing the hidden-layer GEMM: a deliberate trade of idle HBM bandwidth for FP32 work. */
int64_t fbp_plan_cache_per_pair(const fbp_plan* plan);

/* ---- parameter packing: reference pytree leaves <-> [m][P] (leaves: "layers" list of (w (m,out,in),
 *      b (m,out)), fbpinns/networks.py:43-47 vmapped at fbpinns/trainers.py:603-607) ------------ */
int fbp_pack_params(const fbp_plan* plan, int64_t m, const float* const* d_w, const float* const* d_b,
                    float* d_params, void* stream);
int fbp_unpack_params(const fbp_plan* plan, int64_t m, const float* d_params, float* const* d_w,
                      float* const* d_b, void* stream);
/* Per-unit activation parameters of layer `layer` (which = 0: a / c, 1: o) as an [m][out] array <-> packed rows
 * (to_packed != 0: array -> rows).  Only for plans whose activation has such parameters (FBP_ACT_ADAPTIVE_*). */
int fbp_pack_extra(const fbp_plan* plan, int64_t m, int32_t layer, int32_t which, float* d_vec, float* d_params,
                   int32_t to_packed, void* stream);
int32_t fbp_plan_n_extra(const fbp_plan* plan);           /* activation parameters per unit: 0, 1 or 2 */

/* ---- index construction (A2-A4): inside_points / inside_models / get_inputs on the device --- */
/* Phase 1: per-point and per-model inside counts of x against the boxes of `d_models` (NULL = all m).
 *   d_pt_count [n]  = number of selected models containing point i   (inside_ips = count > 0)
 *   d_model_count [n_models] = number of points inside each selected model (inside_ims = count > 0)
 * Comparison is the reference's float32 `x >= xmin & x <= xmax` (fbpinns/decompositions.py:217-227). */
int fbp_inside_count(const float* d_x, int64_t n, int32_t xd, const float* d_sub_static, int32_t m,
                     const int32_t* d_models, int32_t n_models, int32_t* d_pt_count, int32_t* d_model_count,
                     void* stream);
/* Indices of non-zero entries of d_count[n] in ascending order (jnp.arange(n)[mask]); *n_out = how many.
 * d_out must hold n entries.  Synchronises the stream. */
int fbp_nonzero_i32(const int32_t* d_count, int64_t n, int32_t* d_out, int64_t* n_out, void* stream);
/* Row gather: d_dst[i] = d_src[d_idx[i]] for rows of `row_floats` floats (x_batch = x_batch_global[ips]). */
int fbp_gather_rows(const float* d_src, const int32_t* d_idx, int64_t n_idx, int32_t row_floats, float* d_dst,
                    void* stream);

/* Phase 2: begin() counts pairs of x against ALL m boxes and returns s (pairs) and q (rows).
 *   d_pos_of_model [m]: position of each model in all_ims (active first, then fixed), -1 if discarded;
 *   d_pou_of_model [m]: integer partition-of-unity id (static "pou" leaf, fbpinns/decompositions.py:179).
 * emit() writes every array of the takes; sizes: m_take,n_take,p_take,spair_*,pos [s]; np_take [q];
 * row_off [q+1]; pt_row_off [n+1]; sub_off [m_all+1].  Both synchronise the stream. */
int fbp_takes_begin(fbp_takes_builder** b, const float* d_x, int64_t n, int32_t xd, const float* d_sub_static,
                    int32_t m, const int32_t* d_pos_of_model, const int32_t* d_pou_of_model, int32_t m_all,
                    void* stream, int64_t* s_out, int64_t* q_out);
int fbp_takes_emit(fbp_takes_builder* b, int32_t* d_m_take, int32_t* d_n_take, int32_t* d_p_take,
                   int32_t* d_np_take, int32_t* d_row_off, int32_t* d_pt_row_off, int32_t* d_sub_off,
                   int32_t* d_spair_point, int32_t* d_spair_row, int32_t* d_spair_sub, int32_t* d_pos,
                   void* stream);
int fbp_takes_destroy(fbp_takes_builder* b);

/* ---- step path (A5-A7 forward, A9 reverse, A10 Adam) ------------------------------------------ */
/* Denominator jets D = sum_i w_i per row: d_dsum [q][C].  Static between active-set changes
 * (window_fn has no trainable input, fbpinns/decompositions.py:196-199). */
int fbp_window_sums(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_sub_static,
                    float* d_dsum, void* stream);

/* Per-pair numerator jets N_c = d^c(u_i * w_i): norm -> FCN jets -> unnorm -> window jets -> Leibniz.
 * (FBPINN_model_inner under the nested jvp, fbpinns/trainers.py:113-118, 213-247).
 * d_pair_out [s][C*ud] in subdomain-sorted order; d_scratch only for generic plans; d_act_cache optional (NULL). */
int fbp_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                const float* d_sub_static, float* d_pair_out, float* d_scratch, int64_t scratch_floats,
                float* d_act_cache, void* stream);

/* Segment sums + partition-of-unity quotient + /npou (fbpinns/trainers.py:163-170) applied to jets:
 * per row N = sum of its pairs in REFERENCE order, jets of N/D by the quotient rule, summed over the
 * point's rows, divided by npou.  d_ujets [n][C*ud] (points without any pair get 0).
 * d_affine (optional, ud = 1): [n][2*C] jets of A then B of a constraining operator of the form
 * constraining_fn(x, u) = A(x) u + B(x) (fbpinns/problems.py:193-199, 312-318, 381-398); when given, the kernel also
 * applies the Leibniz rule so that d_ujets holds the jets of the CONSTRAINED solution (fbpinns/trainers.py:174). */
int fbp_reduce_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_pair_out,
                       const float* d_dsum, const float* d_affine, float* d_ujets, void* stream);
/* The same in two steps, for the multi-GPU halo exchange: row sums d_nsum [q][C*ud] of the LOCAL pairs,
 * then — after the partial sums of rows shared with other ranks have been added — the quotient rule from row sums.
 * d_out_row (optional): point p is written to row d_out_row[p] of d_ujets, or skipped when that is negative (a sharded host
 * keeps only the points it owns, compactly). */
int fbp_row_sums(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_pair_out, float* d_nsum, void* stream);
int fbp_reduce_rows_forward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_nsum, const float* d_dsum,
                            const float* d_affine, const int32_t* d_out_row, float* d_ujets, void* stream);
/* Transpose of the above: cotangent of ujets [n][C*ud] -> cotangent of row numerators d_grow [q][C*ud]. */
int fbp_reduce_backward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_ujets_bar,
                        const float* d_dsum, const float* d_affine, float* d_grow, void* stream);

/* Reverse pass through the pairs of the ACTIVE subdomains: d_grads [m_active][P] (overwritten when
 * accumulate == 0, added to otherwise — several constraints share one gradient buffer).
 * `accumulate` is a bit set: 1 = add into d_grads; FBP_BWD_NO_REDUCE (2) = run the pair kernels of the view's work
 * list only (per-item partials stay in d_gpart); FBP_BWD_REDUCE_ONLY (4) = only sum the partials into d_grads;
 * FBP_BWD_DIRECT (8, alone) = the caller asserts that the work list has exactly ONE item per active subdomain
 * (d_sub_item_off[i] == i): the tiled / tensor kernels then write the rows of d_grads themselves (no partial buffer, no
 * reduction pass: 58 MB of traffic less at cfg 5); ignored by the generic family.
 * The split lets a multi-GPU host run interior and boundary work items as separate launches around the halo
 * exchange (tiled plans only).
 * d_gpart: workspace of fbp_backward_workspace_floats() floats (split-subdomain partial sums). */
int64_t fbp_backward_workspace_floats(const fbp_plan* plan, const fbp_takes_view* tv);
int fbp_backward(const fbp_plan* plan, const fbp_takes_view* tv, const float* d_x, const float* d_params,
                 const float* d_sub_static, const float* d_grow, float* d_grads, int32_t accumulate,
                 float* d_gpart, float* d_scratch, int64_t scratch_floats, const float* d_act_cache, void* stream);

/* optax.adam + apply_updates (fbpinns/trainers.py:294-295, 430) on rows of [.][P]:
 * for i < n_rows: row r = d_row_ids ? d_row_ids[i] : i of d_params/d_mu/d_nu; gradient row i of d_grads.
 * d_count: device int32 step counter shared by the whole tree; incremented once per call when
 * `increment` != 0 (use increment=0 for the 2nd, 3rd ... buffer updated in the same step). */
int fbp_adam_step(float* d_params, float* d_mu, float* d_nu, const float* d_grads, const int32_t* d_row_ids,
                  int64_t n_rows, int64_t row_len, int32_t* d_count, int32_t increment, float lr, float b1,
                  float b2, float eps, float eps_root, void* stream);

/* FP32 FMA pipe micro-benchmark used as the roofline denominator by bench.py: runs `iters` dependent-free
 * FFMA per thread on a full grid; returns achieved TFLOP/s through *tflops (synchronises). */
int fbp_fma_peak(int32_t iters, float* tflops, void* stream);
/* The same with the packed instruction fma.rn.f32x2 (SASS FFMA2, sm_100+): 16 FMAs per thread per iteration. */
int fbp_ffma2_peak(int32_t iters, float* tflops, void* stream);

/* ---- halo exchange of the sharded step over NVLink peer memory (SURVEY §8b fbp_halo_*, §8e) ---------------------------
 * Replaces, for one node, the reference-side collective a subdomain-sharded FBPINN_model needs around its segment sums
 * (fbpinns/trainers.py:160-170): the partial row sums of points shared across a shard boundary travel as direct stores into
 * the owner's receive buffer, the row cotangents travel back the same way.  Every rank owns one symmetric allocation per
 * constraint:   int32 flags[4][FBP_HALO_MAX_WORLD]  followed by the float receive regions;  `fbp_halo_peers` holds, for every
 * rank, device pointers (valid on the CALLING GPU: cudaIpc / symmetric-memory mappings) to its flag block and data region.
 * Both calls are asynchronous stream work, capturable in CUDA graphs; waits are bounded (trap, never a hang). */
#define FBP_HALO_MAX_WORLD 8
typedef struct fbp_halo_peers {
    float* data[FBP_HALO_MAX_WORLD];      /* base of rank j's receive regions                     */
    int32_t* flags[FBP_HALO_MAX_WORLD];   /* rank j's flags[4][FBP_HALO_MAX_WORLD], zero-initialised */
} fbp_halo_peers;
/* Sends rows d_send_idx[r] of d_rows (row_floats floats each) to the peers.  d_blocks[n_blocks][4] = (peer, first send-list
 * position, end position, number of blocks of that peer): one CTA per entry.  d_dst_off[j] = float offset inside peer j's
 * data region where this rank's rows start (position r of the send list lands at row r there).  dir 0 = forward (partial
 * sums to the owners), 1 = reverse (cotangents back to the sharers); d_epoch[2] = exchange counters of this rank (device),
 * d_ticket[FBP_HALO_MAX_WORLD] zero-initialised scratch. */
int fbp_halo_push(const float* d_rows, int32_t row_floats, const int32_t* d_send_idx, const int32_t* d_blocks, int32_t n_blocks,
                  const fbp_halo_peers* peers, const int64_t* d_dst_off, int32_t me, int32_t dir, const int32_t* d_epoch,
                  int32_t* d_ticket, void* stream);
/* Waits for the peers in `from_mask` and combines what they sent: mode 0: d_rows[d_tgt[i]] += sum of the receive-region rows
 * d_src_pos[d_src_ptr[i] .. d_src_ptr[i+1]) (fixed order: deterministic), mode 1: d_rows[d_tgt[i]] = row i of the receive region (d_src_* unused).
 * Acknowledges to the senders and advances d_epoch[dir].  MUST be called by every rank once per exchange (n_tgt may be 0). */
int fbp_halo_pull(float* d_rows, int32_t row_floats, const int32_t* d_tgt, const int32_t* d_src_ptr, const int32_t* d_src_pos,
                  int32_t n_tgt, const fbp_halo_peers* peers, int64_t my_off, uint32_t from_mask, int32_t me, int32_t world,
                  int32_t dir, int32_t mode, int32_t* d_epoch, int32_t* d_done, void* stream);
/* d_dst (n, row_floats) = scale * d_src[d_inv[row]] where d_inv[row] >= 0, else 0: the owned-row scatter (zero fill + copy +
 * ownership weight) of the sharded reverse pass in one launch. */
int fbp_scatter_rows(const float* d_src, const int32_t* d_inv, int64_t n, int32_t row_floats, float scale, float* d_dst,
                     void* stream);

/* Self-test of the tensor family's MMA form: d_out[128][32] = d_a[128][32] * d_w[32][32]^T computed with the same
 * tcgen05 path the mode-3 kernels use (A written row-wise into tensor memory, B through a shared-memory descriptor,
 * 3xTF32).  `variant` = 0 for the shipped conventions; bits 0-2 probe alternatives (see fbp_tc.cu). */
int fbp_tc_selftest(const float* d_a, const float* d_w, float* d_out, int32_t variant, void* stream);
/* Self-test of the operand form of the tensor-core weight gradient: both operands from shared memory, K-major with a
 * padded K step, contraction over the 128 points of a tile accumulated quarter by quarter:
 * d_out[128][32] = sum_p d_a[p][0..128) (x) d_b[p][0..32), one TF32 pass on TF32-rounded inputs. */
int fbp_tc_selftest_g(const float* d_a, const float* d_b, float* d_out, int32_t variant, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FBPINN_B200_H */
