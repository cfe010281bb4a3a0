// replay_buffer.cu -- device-resident replay memory: minibatch assembly as one gather launch (sm_100a).
//
// The reference keeps its transitions in a host deque (/root/reference/airfoil_dqn.py:48-67, ReplayMemory.push /
// .sample) and collates every minibatch on the host with torch_geometric's DataLoader (:256,268) before moving it
// to the device.  Here the transitions live in fixed-size device slots -- x [cap][n_max][F] f32 and the edge list
// [cap][2][e_max] as slot-local int32 ids, for the state and the next state -- and a sampled minibatch is
// assembled by ONE launch: CTA (b, which) copies the node features of transition idx[b] to rows
// [ptr[g], ptr[g] + n) of the collated matrix, its edges to columns [eptr[g], eptr[g] + E) of the PyG-style int64
// edge_index with the node offset added, and fills the `batch` vector.  The few hundred bytes of per-graph offsets
// come from the host, which mirrors the slots' sizes and therefore knows every launch parameter without
// synchronising; the collated tensors are identical to Batch.from_data_list of the same transitions.
#include "mdq_common.cuh"

namespace {

struct GatherSide {
    const float *x_buf;        // [cap][n_max][F]
    const int *ei_buf;         // [cap][2][e_max], slot-local ids
    const int *n_nodes;        // [cap]
    const int *n_edges;        // [cap]
    const int *gslot;          // [B] row of this side's batch for transition b, or -1 (terminal next state)
    const int *ptr;            // [G + 1] node offsets of this side's batch
    const int *eptr;           // [G + 1]
    float *x_out;              // [sum n][F]
    long long *ei_out;         // [2][E_tot]
    long long *batch_out;      // [sum n]
    long long e_tot;
};

__global__ void __launch_bounds__(256) k_replay_gather(GatherSide s0, GatherSide s1, const long long *__restrict__ idx,
                                                       int n_max, int e_max, int F, const int *__restrict__ act_buf,
                                                       const float *__restrict__ rew_buf, int *__restrict__ act_out,
                                                       float *__restrict__ rew_out)
{
    const GatherSide &s = blockIdx.y ? s1 : s0;
    const int b = blockIdx.x;
    if (blockIdx.y == 0 && threadIdx.x == 0 && act_out) {
        act_out[b] = act_buf[idx[b]];
        rew_out[b] = rew_buf[idx[b]];
    }
    const int g = s.gslot ? s.gslot[b] : b;
    if (g < 0) return;
    const long long slot = idx[b];
    const int n = s.n_nodes[slot], E = s.n_edges[slot];
    const int p0 = s.ptr[g], e0 = s.eptr[g];
    // node features: n*F contiguous floats in the slot -> contiguous in the collated matrix (float4 when aligned)
    const float *xs = s.x_buf + (size_t)slot * n_max * F;
    float *xd = s.x_out + (size_t)p0 * F;
    const int nf = n * F;
    if ((((size_t)p0 * F) & 3) == 0 && (((size_t)n_max * F) & 3) == 0) {
        const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
        float4 *xd4 = reinterpret_cast<float4 *>(xd);
        for (int i = threadIdx.x; i < (nf >> 2); i += blockDim.x) xd4[i] = __ldg(xs4 + i);
        for (int i = (nf & ~3) + threadIdx.x; i < nf; i += blockDim.x) xd[i] = __ldg(xs + i);
    } else {
        for (int i = threadIdx.x; i < nf; i += blockDim.x) xd[i] = __ldg(xs + i);
    }
    const int *es = s.ei_buf + (size_t)slot * 2 * e_max;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        s.ei_out[e0 + e] = (long long)__ldg(es + e) + p0;
        s.ei_out[s.e_tot + e0 + e] = (long long)__ldg(es + e_max + e) + p0;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) s.batch_out[p0 + i] = g;
}

// one transition into its slot: x rows, edge ids (int64 -> slot-local int32), sizes
__global__ void __launch_bounds__(256) k_replay_store(const float *__restrict__ x, int ldx, int n, int F,
                                                      const long long *__restrict__ ei, int E, float *__restrict__ x_slot,
                                                      int *__restrict__ ei_slot, int e_max, int *__restrict__ n_nodes,
                                                      int *__restrict__ n_edges)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * F; i += gridDim.x * blockDim.x) {
        const int r = i / F, f = i - r * F;
        x_slot[i] = x[(size_t)r * ldx + f];
    }
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        ei_slot[e] = (int)ei[e];
        ei_slot[e_max + e] = (int)ei[(size_t)E + e];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *n_nodes = n;
        *n_edges = E;
    }
}

}  // namespace

extern "C" {

int mdq_replay_store(const float *x, int ldx, int n, int F, const int64_t *edge_index, int E, float *x_buf,
                     int32_t *ei_buf, int32_t *n_nodes, int32_t *n_edges, int64_t slot, int n_max, int e_max,
                     void *stream)
{
    if (!x_buf || !ei_buf || !n_nodes || !n_edges || n < 0 || E < 0 || n > n_max || E > e_max || slot < 0 ||
        (n > 0 && !x) || (E > 0 && !edge_index) || ldx < F || F < 1) {
        mdq::set_error("mdq_replay_store: bad argument (n %d / n_max %d, E %d / e_max %d)", n, n_max, E, e_max);
        return MDQ_EINVAL;
    }
    const int work = n * F > E ? n * F : E;
    const int grid = work > 0 ? (work + 255) / 256 : 1;
    k_replay_store<<<grid > 64 ? 64 : grid, 256, 0, (cudaStream_t)stream>>>(
        x, ldx, n, F, reinterpret_cast<const long long *>(edge_index), E, x_buf + (size_t)slot * n_max * F,
        ei_buf + (size_t)slot * 2 * e_max, e_max, n_nodes + slot, n_edges + slot);
    return mdq::check_launch("k_replay_store");
}

int mdq_replay_gather(const float *x_buf, const int32_t *ei_buf, const int32_t *n_nodes, const int32_t *n_edges,
                      const float *xn_buf, const int32_t *ein_buf, const int32_t *nn_nodes, const int32_t *nn_edges,
                      int n_max, int e_max, int F, const int64_t *idx, int B, const int32_t *ptr, const int32_t *eptr,
                      int64_t e_tot, float *x_out, int64_t *ei_out, int64_t *batch_out, const int32_t *next_slot,
                      const int32_t *nptr, const int32_t *neptr, int64_t ne_tot, float *xn_out, int64_t *ein_out,
                      int64_t *nbatch_out, const int32_t *act_buf, const float *rew_buf, int32_t *act_out, float *rew_out,
                      void *stream)
{
    if (!x_buf || !ei_buf || !n_nodes || !n_edges || !idx || !ptr || !eptr || !x_out || !ei_out || !batch_out || B < 1 ||
        n_max < 1 || e_max < 0 || F < 1) {
        mdq::set_error("mdq_replay_gather: bad argument");
        return MDQ_EINVAL;
    }
    if (act_out && (!act_buf || !rew_buf || !rew_out)) {
        mdq::set_error("mdq_replay_gather: action / reward buffers missing");
        return MDQ_EINVAL;
    }
    const bool with_next = xn_out != nullptr;
    if (with_next && (!xn_buf || !ein_buf || !nn_nodes || !nn_edges || !next_slot || !nptr || !neptr || !ein_out || !nbatch_out)) {
        mdq::set_error("mdq_replay_gather: next-state buffers missing");
        return MDQ_EINVAL;
    }
    GatherSide s0{x_buf, ei_buf, n_nodes, n_edges, nullptr, ptr, eptr, x_out, reinterpret_cast<long long *>(ei_out),
                  reinterpret_cast<long long *>(batch_out), e_tot};
    GatherSide s1 = s0;
    if (with_next)
        s1 = GatherSide{xn_buf, ein_buf, nn_nodes, nn_edges, next_slot, nptr, neptr, xn_out,
                        reinterpret_cast<long long *>(ein_out), reinterpret_cast<long long *>(nbatch_out), ne_tot};
    k_replay_gather<<<dim3(B, with_next ? 2 : 1), 256, 0, (cudaStream_t)stream>>>(
        s0, s1, reinterpret_cast<const long long *>(idx), n_max, e_max, F, act_buf, rew_buf, act_out, rew_out);
    return mdq::check_launch("k_replay_gather");
}

}  // extern "C"
