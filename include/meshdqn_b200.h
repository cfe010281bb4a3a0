/*
 * meshdqn_b200.h -- C ABI of libmeshdqn_b200.so (hand-written sm_100a CUDA kernels).
 *
 * The reference (BaratiLab/MeshDQN) is pure Python with no FFI; its "plugin
 * boundary" for the hot path is the Python surface of airfoilgcnn.py and
 * Env2DAirfoil.py, whose arithmetic lives in torch_geometric / DOLFIN / shapely.
 * Each entry point below replaces one of those third-party call sites; the
 * Python host mirror (meshdqn_b200/*.py) keeps the reference's class and method
 * names and calls these through ctypes.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer owned by the caller unless its name starts
 *     with h_ (host); the callee never allocates, frees or retains them;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), no internal
 *     synchronisation, re-entrant per stream;
 *   - return 0 on success, a negative MDQ_E* code on failure (never throws);
 *     mdq_last_error() gives a thread-local message;
 *   - one CUDA device per process (the deployment model: one process per GPU): the kernels' opt-in shared-memory
 *     attributes are set once per process; several host threads may call concurrently on their own streams.
 */
#ifndef MESHDQN_B200_H
#define MESHDQN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDQ_OK 0
#define MDQ_EINVAL (-1)   /* bad argument / unsupported size */
#define MDQ_ECUDA (-2)    /* CUDA runtime error (see mdq_last_error) */
#define MDQ_ESMEM (-3)    /* problem does not fit the fused kernel's shared memory */

#define MDQ_MAX_BLOCKS 6
#define MDQ_BLOCK_SAGE 0
#define MDQ_BLOCK_GCN 1

const char *mdq_last_error(void);
int mdq_version(void);
/* number of kernel launches issued through this library by the calling process */
int64_t mdq_launch_count(void);
/* a replayed CUDA graph launches the kernels recorded at capture time without passing through this library: the host
 * mirror adds that number per replay so that mdq_launch_count() stays the number of kernel launches that ran */
void mdq_launch_count_add(int64_t n);

/* ------------------------------------------------------------------------------------
 * Q-network (replaces torch_geometric SAGEConv/GCNConv/TopKPooling/global pools +
 * nn.Linear/softmax/argmax in /root/reference/airfoilgcnn.py:85-145 and :170-209).
 *
 * All parameters live in ONE flat fp32 buffer; offsets are in floats.  Dense weights are
 * stored transposed ([in][out], out contiguous) so a warp reads consecutive outputs:
 *   SAGE block: w_off -> [2*kin][width]  rows 0..kin-1 = lin_l.weight^T, kin..2kin-1 = lin_r.weight^T
 *               b_off -> lin_l.bias [width]
 *   GCN  block: w_off -> [kin][width] = lin.weight^T ; b_off -> bias [width]
 *   pool_off -> TopKPooling.weight [width]
 *   lin_off[i] -> [lin_in[i]][lin_out[i]] = lin{i+1}.weight^T ; lin_boff[i] -> bias
 * ------------------------------------------------------------------------------------ */
typedef struct {
    int32_t type;      /* MDQ_BLOCK_SAGE | MDQ_BLOCK_GCN */
    int32_t kin;       /* input features of this block */
    int32_t w_off, b_off, pool_off;
} mdq_block_t;

typedef struct {
    int32_t n_blocks;
    int32_t width;        /* conv_width (multiple of 4, <= 256) */
    int32_t in_dim;       /* features the first block consumes */
    int32_t in_col0;      /* first used column of data.x (AirfoilGCNN: x[:, [2,3]] -> 2) */
    int32_t x_stride;     /* columns of data.x */
    float ratio;          /* TopKPooling ratio */
    int32_t softmax;      /* 1: softmax over the last layer (NodeRemovalNet), 0: raw (AirfoilGCNN) */
    int32_t out_dim;
    int32_t lin_off[3], lin_boff[3], lin_in[3], lin_out[3];
    int32_t n_params;     /* floats in the flat buffer */
    mdq_block_t blk[MDQ_MAX_BLOCKS];
} mdq_net_t;

/* Shared memory (bytes) the fused kernels need for graphs of at most max_n nodes / max_e edges
 * with `gpc` graphs per CTA; backward != 0 for the recompute+backward kernel.  Host-callable, no GPU. */
int64_t mdq_qnet_smem_bytes(const mdq_net_t *net, int max_n, int max_e, int gpc, int backward);
/* resident CTAs per SM of the forward (backward != 0: backward) kernel for these sizes (needs a device) */
int mdq_qnet_occupancy(const mdq_net_t *net, int max_n, int max_e, int backward);
/* graphs per CTA the forward launch will use for this batch (host-callable) */
int mdq_qnet_pick_gpc(const mdq_net_t *net, int n_graphs, int max_n, int max_e);

/* Forward: one CTA per `gpc` graphs.
 *   x [sum_n, x_stride] f32; edge_src/edge_dst [sum_e] i64 GLOBAL node ids (PyG edge_index rows 0/1);
 *   node_ptr / edge_ptr [B+1] i32.  out [B, out_dim] f32 (softmax "Q-values");
 *   embedding (nullable) [B, 2*width]; argmax (nullable) [B] i32 (first maximum). */
int mdq_qnet_forward(const mdq_net_t *net, const float *params, const float *x, const int64_t *edge_src,
                     const int64_t *edge_dst, const int32_t *node_ptr, const int32_t *edge_ptr, int n_graphs,
                     int max_n, int max_e, float *out, float *embedding, int32_t *argmax, void *stream);

/* Profiling aid: when set (device int64[512], NULL to disable), CTA 0 of the next qnet launches writes clock64()
 * at its phase boundaries (MDQ_TRACE in csrc/gnn_fused.cu: entries 0..127; STG_TRACE in csrc/gnn_staged.cuh: stage 0 at
 * 0.., stage 1 at 32.., stage 2 at 96.., backward 1 at 256..). */
void mdq_qnet_set_trace(int64_t *device_buf);

/* Rows of (delta, input) pairs the backward kernel emits for the weight-gradient pass. */
int64_t mdq_qnet_bwd_workspace_floats(const mdq_net_t *net, int n_graphs, int max_n);

/* Backward (recomputes the forward per graph, deterministic, atomics-free):
 *   grad_out [B, out_dim] = dL/d out.  Writes the flat gradient grad [n_params] (overwritten).
 *   workspace: mdq_qnet_bwd_workspace_floats() floats. */
int mdq_qnet_backward(const mdq_net_t *net, const float *params, const float *x, const int64_t *edge_src,
                      const int64_t *edge_dst, const int32_t *node_ptr, const int32_t *edge_ptr, int n_graphs,
                      int max_n, int max_e, const float *grad_out, float *grad, float *workspace, void *stream);

/* Fused replay gradient (airfoil_dqn.py:240-310): the backward kernel evaluates the Huber term itself from its
 * recomputed forward, so the selected net's forward is never launched separately.
 *   mode 1 (select): the batch is the B states; index = next_slot [B]; q_other = Q2(s') [n_next, A] (may be NULL
 *                    when every transition is terminal); scalar [B] receives pred_b = Q1(s_b)[a_b].
 *   mode 2         : the batch is the n_next next-states; index = owner [n_next] (transition of each row);
 *                    q_other = Q1(s) [B, A]; scalar [n_next] receives max_a Q2(s'_g).
 * next_slot [B] is needed in both modes for the loss; loss = mean Huber over `batch` transitions; grad [n_params]. */
int mdq_qnet_replay_backward(const mdq_net_t *net, const float *params, const float *x, const int64_t *edge_src,
                             const int64_t *edge_dst, const int32_t *node_ptr, const int32_t *edge_ptr, int n_graphs,
                             int max_n, int max_e, int mode, const int32_t *action, const float *reward,
                             const int32_t *index, const int32_t *next_slot, const float *q_other, int batch, float gamma,
                             float *scalar, float *loss, float *grad, float *workspace, void *stream);

/* ------------------------------------------------------------------------------------
 * Staged tensor-core path (csrc/gnn_staged.cuh) for NodeRemovalNet-shaped networks (SAGE, SAGE, GCN, GCN, width 128,
 * MLP 256-128-64-out) on batches of state graphs with <= 256 nodes: the same function as mdq_qnet_forward /
 * mdq_qnet_backward / mdq_qnet_replay_backward (same arguments, same meaning), computed as four launches whose node /
 * MLP GEMMs run on tcgen05 as 3xTF32 (fp32 operands split hi + lo, fp32 accumulation in TMEM): Q within ~1e-6 of the
 * fp32 kernels instead of bit-equal to them.  Replaces the same torch_geometric / nn.Linear call sites
 * (/root/reference/airfoilgcnn.py:94-143, airfoil_dqn.py:258-305).
 *   wsplit: mdq_qnet_staged_wsplit_floats() floats, filled by mdq_qnet_staged_wsplit() from the flat parameters and
 *           refreshed by the caller after every weight update (one launch);
 *   workspace: mdq_qnet_staged_workspace_floats() floats, 16-byte aligned, one per concurrent stream;
 *   edge_src / edge_dst: the two rows of edge_index as int64 (PyG) or, with edge_i32 != 0, as int32 (the replay
 *           minibatch travels over PCIe with 32-bit edges).
 * ------------------------------------------------------------------------------------ */
int mdq_qnet_staged_supported(const mdq_net_t *net, int max_n, int max_e);
int64_t mdq_qnet_staged_wsplit_floats(const mdq_net_t *net);
int mdq_qnet_staged_wsplit(const mdq_net_t *net, const float *params, float *wsplit, void *stream);
int64_t mdq_qnet_staged_workspace_floats(const mdq_net_t *net, int n_graphs, int max_n, int max_e, int backward);
int mdq_qnet_staged_forward(const mdq_net_t *net, const float *params, const float *wsplit, const float *x,
                            const void *edge_src, const void *edge_dst, int edge_i32, const int32_t *node_ptr,
                            const int32_t *edge_ptr, int n_graphs, int max_n, int max_e, float *out, float *embedding,
                            int32_t *argmax, float *workspace, void *stream);
int mdq_qnet_staged_backward(const mdq_net_t *net, const float *params, const float *wsplit, const float *x,
                             const void *edge_src, const void *edge_dst, int edge_i32, const int32_t *node_ptr,
                             const int32_t *edge_ptr, int n_graphs, int max_n, int max_e, const float *grad_out, float *grad,
                             float *workspace, void *stream);
int mdq_qnet_staged_replay_backward(const mdq_net_t *net, const float *params, const float *wsplit, const float *x,
                                    const void *edge_src, const void *edge_dst, int edge_i32, const int32_t *node_ptr,
                                    const int32_t *edge_ptr, int n_graphs, int max_n, int max_e, int mode,
                                    const int32_t *action, const float *reward, const int32_t *index,
                                    const int32_t *next_slot, const float *q_other, int batch, float gamma, float *scalar,
                                    float *loss, float *grad, float *workspace, int phase, uint32_t *tail_sync,
                                    void *stream);
/* mdq_qnet_staged_replay_backward differs from mdq_qnet_replay_backward in two ways: the flat gradient's entries of
 * blocks the forward never uses are left untouched (keep a persistent zero-initialised buffer), and `phase` splits the
 * call -- 0: everything; 1: stages 0 / 1 of the selected net only (they do not read q_other, so they may be enqueued on
 * a second stream beside the other net's forward); 2: the remaining launches (same arguments, same workspace);
 * 3: stages 0 / 1 + the tail backward; 4: backward 1 + weight gradients (3 then 4 == 1 then 2 == 0).
 * tail_sync (phase 3 only, nullable): three zero-initialised device words.  With it the tail kernel may be enqueued BEFORE
 * q_other has been computed: it runs its forward part and waits, on the device, for one mdq_stream_post(tail_sync, s)
 * enqueued behind the other net's forward on ANOTHER stream, exactly where the loss first reads q_other (posts and tail
 * launches pair up one to one, in order).  Never enqueue the post behind the waiting call on the same stream, and not
 * under a tool that serialises kernels (ncu, compute-sanitizer): the wait traps after ~2 s instead of hanging. */
int mdq_stream_post(uint32_t *tail_sync, void *stream);

/* ------------------------------------------------------------------------------------
 * Layered forward for ONE large graph (a state graph that does not fit the fused kernel's shared memory, e.g. the
 * ~0.5M-node graph of a ~1M-triangle mesh).  Same network, same results (to fp32 rounding) as mdq_qnet_forward:
 * CSR message passing (warp per row), node GEMMs, radix-sort TopK, ordered edge filter, readout, MLP head.
 *   gemm_mode 0: fp32 FFMA node GEMMs;  1: tcgen05 3xTF32 (tensor cores, TMEM accumulator; conv width 128).
 *   wsplit (gemm_mode 1): 3 * n_params floats; the conv weight at w_off is stored at 3 * w_off as
 *     hi [Kpad/4][16][8][4] then lo (same shape), hi = w & 0xffffe000, lo = w - hi,
 *     element [c][g][r][kk] = W'[4c+kk][8g+r]  (K-major core matrices of the UMMA canonical layout), where W' is W with
 *     its rows moved to the A operand's columns: GCN blocks W' = W (Kpad = K rounded up to 8); SAGE blocks
 *     W'[0:F] = lin_r^T, W'[Fp:Fp+F] = lin_l^T, other rows zero (Fp = 4*ceil(F/4), Kpad = 2 Fp rounded up to 8).
 * ------------------------------------------------------------------------------------ */
int64_t mdq_qnet_layered_workspace_bytes(const mdq_net_t *net, int n_nodes, int n_edges);
int mdq_qnet_forward_layered(const mdq_net_t *net, const float *params, const float *wsplit, const float *x,
                             const int64_t *edge_src, const int64_t *edge_dst, int n_nodes, int n_edges, int gemm_mode,
                             float *out, float *embedding, int32_t *argmax, void *workspace, int64_t workspace_bytes,
                             void *stream);

/* Building blocks of the layered path, exported for parity tests and roofline measurement.
 * mdq_csr_build: CSR by destination with rows in edge order; ecount is a DEVICE int (<= ecap);
 *   scratch: mdq_csr_build_scratch_words(ecap, n) int32 words.
 * mdq_sage_aggregate (torch_geometric SAGEConv message passing, airfoilgcnn.py:94,100):
 *   A[i] = [ x[i, col0:col0+F] (zero-padded to Fp = 4*ceil(F/4)) | mean_{j->i} x[j, col0:col0+F] (padded to Fp) | 0 ... ]
 *   with row stride lda >= 2 Fp, lda % 4 == 0, A 16-byte aligned.
 * mdq_node_gemm: C[M,N] = epilogue(A[rows][K] . W[K][N]): + bias, ReLU, score[m] = tanh(h.pool/||pool||),
 *   stored row scaled by row_scale[source row]; every pointer after W may be NULL. */
int mdq_csr_build(const int32_t *src, const int32_t *dst, const int32_t *ecount, int ecap, int n, int32_t *row_ptr,
                  int32_t *col, int32_t *scratch, void *stream);
int64_t mdq_csr_build_scratch_words(int ecap, int n);
int mdq_sage_aggregate(const float *x, int ldx, int col0, int F, const int32_t *row_ptr, const int32_t *col, int n,
                       float *A, int lda, void *stream);
int mdq_node_gemm(const float *A, const int32_t *rows, int lda, int K, int M, int N, const float *W, const float *wsplit,
                  const float *bias, const float *pool, const float *row_scale, int relu, int gemm_mode, float *C,
                  float *score, void *stream);

/* Replay-minibatch loss (replaces /root/reference/airfoil_dqn.py:264,267-283,303-304):
 *   pred_b = q1[b, action[b]];  target_b = reward[b] + gamma * (nonfinal[b] ? max_a q2[slot[b], a] : 0)
 *   loss = mean_b huber(pred_b - target_b, delta = 1)
 * select != 0: grad_q1 [B, A] = dloss/dq1 (grad_q2 untouched); select == 0: grad_q2 [B2, A] = dloss/dq2.
 * next_slot[b] = row of q2 holding b's next state, or -1 when the transition is terminal. */
int mdq_huber_replay(const float *q1, const float *q2, const int32_t *action, const float *reward,
                     const int32_t *next_slot, int batch, int n_next, int out_dim, float gamma, int select,
                     float *loss, float *grad_q1, float *grad_q2, void *stream);

/* Adam step on flat buffers with torch.optim.Adam semantics (L2 weight decay added to the gradient;
 * /root/reference/airfoil_dqn.py:172-173).  grad is multiplied by grad_scale first (1/world_size after
 * an all-reduce sum).  step is the 1-based step count. */
int mdq_adam_step(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, float grad_scale, int step,
                  void *stream);

/* Same update with the step count kept on the device: step_dev is TWO int32, [0] = steps taken so far (advanced by the
 * kernel's last block), [1] = block ticket, both zero before the first call.  No host-computed bias corrections in the
 * launch arguments, so a captured CUDA graph of the training step stays valid. */
int mdq_adam_step_dev(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, float grad_scale, int32_t *step_dev,
                      void *stream);

/* Fused gradient all-reduce + Adam over NVLink peer memory: ONE launch replaces the NCCL all-reduce of the flat
 * gradient (/root/reference/airfoil_dqn.py:326-336 ships it through the Ray object store) and the optimizer step.
 *   h_peer_stage (HOST array [world]): device addresses of every rank's stage buffer as mapped in THIS process
 *     (peer memory, 16-byte aligned, mdq_allreduce_stage_floats(n, world) floats each, zero before the first call);
 *   grad: this rank's gradient (sum over its own transitions); the update uses mean over ranks = sum / world, summed in
 *     rank order on every rank (bit-identical weights everywhere);
 *   step_dev: as mdq_adam_step_dev; block_counter: TWO device uint32, zero before the first call.
 * Two-shot exchange of posted stores: gradient slices to their owners, reduced slices back to everyone (2 (world-1)/world
 * x 4n bytes per rank over NVLink); every wait polls local memory. */
int64_t mdq_allreduce_stage_floats(int64_t n, int world);
int mdq_allreduce_adam(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                       float beta1, float beta2, float eps, float weight_decay, int32_t *step_dev,
                       const uint64_t *h_peer_stage, int rank, int world, uint32_t *block_counter, void *stream);

/* ------------------------------------------------------------------------------------
 * Environment step, float64 geometry (replaces DOLFIN / shapely calls in
 * /root/reference/Env2DAirfoil.py and flow_solver.py).  Coordinates are [n,2] f64 row-major,
 * cells [n,3] i32 with ascending vertex ids per cell (DOLFIN's ordering after close()).
 * ------------------------------------------------------------------------------------ */

/* Exclusive scan of int32 in[0..n) -> out[0..n] (out[n] = total); one CTA, stream-ordered. */
int mdq_scan_i32(const int32_t *in, int32_t *out, int n, void *stream);

/* Mesh topology (replaces Mesh.init / BoundaryMesh, flow_solver.py:75,247; Env2DAirfoil.py:464).
 * Deterministic and sort-free: per-vertex sorted neighbour lists give lexicographic edge ids.
 *   nbr_ptr [nv+1], nbr_idx [6*nc] ascending neighbours; vc_ptr [nv+1], vc_idx [3*nc] ascending cells;
 *   edge_base [nv+1] (edges owned by their lower endpoint), edges [3*nc,2] (first ne rows valid),
 *   cell_edges [nc,3] (edge opposite local vertex i), edge_ncells [3*nc], edge_cell [3*nc] = 4*cell+local for
 *   exterior facets, on_boundary [nv] u8, bverts [nv] (ascending boundary vertex ids),
 *   counts int32[4]: [0] = ne, [1] = #boundary vertices, [3] = 1 if a vertex exceeded the neighbour capacity.
 *   scratch: int32 [2*nv + 2]. */
int mdq_mesh_topology(const int32_t *cells, int nc, int nv, int32_t *nbr_ptr, int32_t *nbr_idx, int32_t *vc_ptr,
                      int32_t *vc_idx, int32_t *edge_base, int32_t *edges, int32_t *cell_edges, int32_t *edge_ncells,
                      int32_t *edge_cell, uint8_t *on_boundary, int32_t *bverts, int32_t *counts, int32_t *scratch,
                      void *stream);

/* Mesh.smooth(iters) (flow_solver.py:67,237): in-place Gauss-Seidel sweep in vertex order, boundary fixed.
 * The exact sequential result is kept by level-scheduling the vertex dependency DAG inside one CTA (8 lanes per
 * vertex, coordinates + adjacency in shared memory); the ~depth*iters dependent updates bound its latency.
 * status: device int32, set to 0 on completion. */
int mdq_mesh_smooth(double *coords, int nv, int nc, const int32_t *nbr_ptr, const int32_t *nbr_idx,
                    const int32_t *vc_ptr, const int32_t *vc_idx, const int32_t *cells, const uint8_t *on_boundary,
                    int iters, int32_t *status, void *stream);

/* Diagnostics only (tools/smooth_bench.py): clock64 / %globaltimer stamps of the last mdq_mesh_smooth launch --
 * [0] kernel start, [1] sweep start, [2] sweep end (cycles); [4], [5] sweep start / end (ns); [6] level count. */
int mdq_debug_smooth_trace(long long *out8);

/* Diagnostics only (tests/test_env_gpu.py): the branch-free division / square-root sequences of the smoothing sweep
 * against the compiler's `/` and sqrt() on n pseudo-random operand pairs (mode 0: any finite bit pattern, mode 1:
 * exponents within +-40 of 1.0).  counts4 (device): [0] accepted quotients that differ, [1] accepted roots that differ,
 * [2], [3] operands the sequences declined (k_smooth redoes those with the ordinary operators). */
int mdq_debug_fast_math_check(uint64_t seed, int64_t n, int mode, uint64_t *counts4, void *stream);

/* FlowSolver.mark_boundaries (flow_solver.py:9-30,194-226): tags [ne] i32 (4 default, 0 walls, 1 airfoil,
 * 2 inflow, 3 outflow) and the removable mask (flow_solver.py:75-78,247-250; numpy `coord not in B`) [nv] u8. */
int mdq_mesh_tags_removable(const double *coords, int nv, const int32_t *edges, const int32_t *edge_ncells, int ne,
                            const int32_t *bverts, int nb, int32_t *tags, uint8_t *removable, void *stream);

/* shapely Polygon.distance(Point) (Env2DAirfoil.py:232,240-241): dist[i] for points coords[idx[i]]
 * (idx nullable = identity) against the closed ring [nr,2]. */
int mdq_polygon_distance(const double *coords, const int32_t *idx, int np, const double *ring, int nr, double *dist,
                         void *stream);

/* Uniform-grid index over the SOURCE mesh M0 for point location.  h_grid (HOST, 6 doubles):
 * x0, y0, 1/dx, 1/dy, gx, gy.  mdq_grid_count -> bin_cnt [gx*gy+1] (entries per bin); the caller scans it
 * (mdq_scan_i32) into bin_ptr and sizes bin_cells = bin_ptr[gx*gy]; mdq_grid_fill writes the cell lists. */
int mdq_grid_count(const double *coords0, const int32_t *cells0, int nc0, const double *h_grid, int32_t *bin_cnt,
                   void *stream);
int mdq_grid_fill(const double *coords0, const int32_t *cells0, int nc0, const double *h_grid,
                  const int32_t *bin_ptr, int32_t *bin_cursor, int32_t *bin_cells, void *stream);

/* Function.interpolate for T snapshots (Env2DAirfoil.py:556-568, :515-522):
 *   target P2 dof points = vertices [nv] then edge midpoints 0.5*a+0.5*b [ne];
 *   located in M0: lowest-index cell with min barycentric >= -tol, else the closest cell (lowest index on ties);
 *   U0 [T][nv0+ne0][2], P0 [T][nv0]  ->  U [T][nv+ne][2], P [T][nv], cell_of [nv+ne] i32.
 *   miss_count: device int32 (points that needed the closest-cell fallback); miss_list [nv+ne] scratch.
 *   coords must be 16-byte aligned and edges 8-byte aligned (vector loads). */
int mdq_interpolate(const double *coords, int nv, const int32_t *edges, int ne, const double *coords0,
                    const int32_t *cells0, const int32_t *cell_edges0, int nv0, int ne0, int nc0,
                    const double *h_grid, const int32_t *bin_ptr, const int32_t *bin_cells, double tol, int T,
                    const double *U0, const double *P0, double *U, double *P, int32_t *cell_of,
                    int32_t *miss_count, int32_t *miss_list, void *stream);

/* Strict mode of the interpolation pin (the reference's "INTERPOLATION BROKE" branch,
 * Env2DAirfoil.py:569-573): after mdq_interpolate / mdq_interpolate_tiled, max_d2 (device f64) = the largest squared
 * distance between a target point that fell outside every source cell and the closest cell it was given
 * (0 when nothing missed).  The host compares it with a macroscopic threshold and returns code 2 above it. */
int mdq_interp_miss_distance(const double *coords, int nv, const int32_t *edges, int ne, const double *coords0,
                             const int32_t *cells0, const int32_t *cell_of, const int32_t *miss_count,
                             const int32_t *miss_list, double *max_d2, void *stream);

/* Tiled re-interpolation for large source meshes (same results as mdq_interpolate, bit for bit).
 * The index over M0 is built once on the host (meshdqn_b200/tile_index.py): k-d leaves of ~256 cells whose whole
 * working set (local coordinates, all T snapshots' P2/P1 coefficients, cell->dof table, micro-grid) is contiguous,
 * so one CTA stages a leaf in shared memory with bulk (TMA) copies and serves every target point inside it. */
typedef struct {
    int32_t n_leaves, depth, T;
    int32_t max_nv, max_np2, max_nc, max_nbin, max_nent; /* padded per-leaf maxima (informational) */
    int64_t u_stride, p_stride;                          /* rows of UL / PL per snapshot */
    const double *tree;         /* [n_leaves-1] split planes, heap order, split axis in the mantissa LSB */
    const int32_t *leaf_info;   /* [n_leaves][16] */
    const double *leaf_rect;    /* [n_leaves][4] micro-grid x0, y0, 1/dx, 1/dy */
    const double *coordsL;      /* [sum nv][2] */
    const double *UL;           /* [T][u_stride][2] */
    const double *PL;           /* [T][p_stride] */
    const int32_t *gidL;        /* [sum nc] global cell ids, ascending per leaf */
    const uint16_t *cvL;        /* [sum nc][6] local dof ids (3 vertices, 3 edges) */
    const uint16_t *binptrL;    /* micro-grid CSR pointers */
    const uint16_t *binsL;      /* micro-grid candidate lists (local cell ids, ascending) */
    const int32_t *leaf_base;   /* [n_leaves+1] target-record bucket offsets (capacity prefix sums) */
    int64_t total_cap;          /* = leaf_base[n_leaves] */
    int32_t smem_bytes;         /* dynamic shared memory per CTA: sized for the typical leaf so that several CTAs share an
                                 * SM; a leaf whose own sections need more is served from HBM by the same CTA
                                 * (0: size for the largest leaf, from the maxima above) */
    int32_t reserved;
} mdq_tile_index_t;

/* Workspace of mdq_interpolate_tiled, in int32 words:
 *   counters: persistent, MUST be zero before the first call; every successful call leaves it zeroed again;
 *   scratch : per call (16-byte aligned), for np = nv + ne target points. */
int64_t mdq_interp_tiled_counter_words(const mdq_tile_index_t *idx);
int64_t mdq_interp_tiled_scratch_words(const mdq_tile_index_t *idx, int np);
/* shared memory (bytes) one CTA of the tiled kernel uses; host-callable */
int64_t mdq_interp_tiled_smem_bytes(const mdq_tile_index_t *idx);

/* Same contract as mdq_interpolate (targets, outputs, miss list); coords0/cells0/cell_edges0/U0/P0 are only
 * touched by the closest-cell fallback for points outside every source cell. */
int mdq_interpolate_tiled(const double *coords, int nv, const int32_t *edges, int ne, const mdq_tile_index_t *idx,
                          const double *coords0, const int32_t *cells0, const int32_t *cell_edges0, int nv0, int ne0,
                          int nc0, const double *U0, const double *P0, double tol, double *U, double *P,
                          int32_t *cell_of, int32_t *miss_count, int32_t *miss_list, int32_t *counters,
                          int32_t *scratch, void *stream);

/* DragProbe/LiftProbe.sample for T <= 8 snapshots (probes.py:23-31,43-50): sum over exterior facets with
 * tag == 1 of |f| (sigma(m_f) n).e_x / e_y.  drag_lift [2][T] f64.  Deterministic fixed-shape reduction. */
int mdq_drag_lift(const double *coords, const int32_t *cells, const int32_t *cell_edges, int nv, int ne,
                  const int32_t *tags, const int32_t *edge_cell, int T, const double *U, const double *P, double mu,
                  double *drag_lift, void *stream);

/* mdq_interpolate followed by mdq_drag_lift on the TARGET mesh in the same two launches: the last block of the miss pass
 * integrates the airfoil facets (tags == 1) over the freshly interpolated U / P -- "interpolation ... followed by a fused
 * surface-integral drag/lift reduction".  Same reduction shape as mdq_drag_lift: bit-identical drag / lift.
 *   cells, cell_edges, tags, edge_cell: of the target mesh (mdq_mesh_topology / mdq_mesh_tags_removable);
 *   ticket: one device uint32, zero before the first call (left zero by every call). */
int mdq_interpolate_drag_lift(const double *coords, int nv, const int32_t *edges, int ne, const double *coords0,
                              const int32_t *cells0, const int32_t *cell_edges0, int nv0, int ne0, int nc0,
                              const double *h_grid, const int32_t *bin_ptr, const int32_t *bin_cells, double tol, int T,
                              const double *U0, const double *P0, double *U, double *P, int32_t *cell_of,
                              int32_t *miss_count, int32_t *miss_list, const int32_t *cells, const int32_t *cell_edges,
                              const int32_t *tags, const int32_t *edge_cell, double mu, double *drag_lift, uint32_t *ticket,
                              void *stream);

/* Env2DAirfoil.get_state (Env2DAirfoil.py:244-315) on device, quirks B1-B3 included:
 *   dist [nrem] f64 distances of the removable vertices (list order), removable_idx [nrem] i32;
 *   picks order[offset : offset+N] of the stable ascending argsort -> n_closest [N] (positions in the
 *   removable list, -1 when out of vertices), coord_map [N], inv_map [nv]; x [N, 3T+2] f32;
 *   edge_index i64 [2][ecap] (PyG layout, first n_edges columns valid). */
int mdq_build_state(const double *dist, const int32_t *removable_idx, int nrem, int offset, int N,
                    const double *coords, int nv, const int32_t *cells, int nc, int T, const double *U, int np2,
                    const double *P, int32_t *n_closest, int32_t *coord_map, int32_t *inv_map, float *x,
                    int64_t *edge_index, int ecap, int32_t *n_edges, void *stream);

/* ---- device-resident replay memory (replaces the host deque + DataLoader collation of airfoil_dqn.py:48-67,256,268) ----
 * Slots: x_buf [cap][n_max][F] f32, ei_buf [cap][2][e_max] i32 (slot-local node ids), n_nodes / n_edges [cap] i32;
 * one set for the states, one for the next states.
 * mdq_replay_store: copy one graph (x [n][ldx], PyG int64 edge_index [2][E]) into `slot`.
 * mdq_replay_gather: assemble the minibatch idx[0..B) as PyG-collated tensors in ONE launch:
 *   x_out [ptr[B]][F], ei_out int64 [2][e_tot] (node offsets added), batch_out int64 [ptr[B]]; ptr / eptr i32 [B+1]
 *   are the offsets of the collated batch (the host mirrors the slot sizes, so it provides them without a sync);
 *   next states (xn_out != NULL): transition b's next state goes to row next_slot[b] (>= 0) of the second batch
 *   with offsets nptr / neptr, terminal transitions (next_slot < 0) are skipped;
 *   act_out / rew_out (nullable) [B] = act_buf / rew_buf [cap] gathered through idx. */
int mdq_replay_store(const float *x, int ldx, int n, int F, const int64_t *edge_index, int E, float *x_buf,
                     int32_t *ei_buf, int32_t *n_nodes, int32_t *n_edges, int64_t slot, int n_max, int e_max,
                     void *stream);
int mdq_replay_gather(const float *x_buf, const int32_t *ei_buf, const int32_t *n_nodes, const int32_t *n_edges,
                      const float *xn_buf, const int32_t *ein_buf, const int32_t *nn_nodes, const int32_t *nn_edges,
                      int n_max, int e_max, int F, const int64_t *idx, int B, const int32_t *ptr, const int32_t *eptr,
                      int64_t e_tot, float *x_out, int64_t *ei_out, int64_t *batch_out, const int32_t *next_slot,
                      const int32_t *nptr, const int32_t *neptr, int64_t ne_tot, float *xn_out, int64_t *ein_out,
                      int64_t *nbatch_out, const int32_t *act_buf, const float *rew_buf, int32_t *act_out, float *rew_out,
                      void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MESHDQN_B200_H */
