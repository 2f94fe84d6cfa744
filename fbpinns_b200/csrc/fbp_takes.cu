// Device-side index construction: inside_points / inside_models
// (fbpinns/decompositions.py:201-227, fbpinns/decompositions_base.py:23-82) and the pair / row bookkeeping
// of get_inputs (fbpinns/trainers.py:332-391), producing arrays that are BIT-EXACT with the reference's:
//   n_take, m_take : row-major non-zeros of the dense (n, m) inside mask (sorted by point, then model),
//                    m_take already re-indexed to positions in all_ims (trainers.py:370-372)
//   p_take, np_take: inverse / first column of jnp.unique(stack[n_take, pou[m_take]], axis=0)
// plus the subdomain-sorted view the kernels consume (a stable counting of pairs by subdomain position).
//
// The inside test is the reference's dense O(n*m) float32 comparison, evaluated with one thread per point and
// the boxes streamed through shared memory — integer/compare work bound by the issue rate, not HBM
// (x is read once, boxes m*2*xd floats per CTA).  Exclusive scans and the stable key sort use CUB (bundled
// with the CUDA toolkit); they run only when the active set changes, never on the step path.
#include "fbp_common.cuh"

#include <cub/cub.cuh>
#include <vector>

namespace {

constexpr int TT = 256;
constexpr int BOX_CHUNK = 1024;   // boxes staged per shared-memory pass (1024 * 2*3 floats = 24 KB)

// lo/hi of box b staged as s_box[(2*d)*BOX_CHUNK + b], s_box[(2*d+1)*BOX_CHUNK + b]
template <int XD>
__device__ __forceinline__ bool inside_box(const float* s_box, int b, const float* xx) {
    bool in = true;
#pragma unroll
    for (int d = 0; d < XD; ++d) {
        float lo = s_box[(2 * d) * BOX_CHUNK + b], hi = s_box[(2 * d + 1) * BOX_CHUNK + b];
        in = in && (xx[d] >= lo) && (xx[d] <= hi);
    }
    return in;
}

// A model whose position is < 0 (not in all_ims: discarded, or owned by another rank) is staged as an empty box
// (lo = +inf, hi = -inf) so that it never matches.
template <int XD>
__device__ __forceinline__ void stage_boxes(float* s_box, int32_t* s_aux, const float* __restrict__ sub_static,
                                            const int32_t* __restrict__ models, const int32_t* __restrict__ aux,
                                            const int32_t* __restrict__ pos, int b0, int nb, int ss) {
    for (int t = threadIdx.x; t < nb; t += blockDim.x) {
        int im = models ? models[b0 + t] : b0 + t;
        const float* rec = sub_static + (int64_t)im * ss;
        const bool keep = pos ? pos[im] >= 0 : true;
#pragma unroll
        for (int d = 0; d < XD; ++d) {
            s_box[(2 * d) * BOX_CHUNK + t] = keep ? rec[d] : INFINITY;
            s_box[(2 * d + 1) * BOX_CHUNK + t] = keep ? rec[XD + d] : -INFINITY;
        }
        if (s_aux) s_aux[t] = aux ? aux[im] : 0;
    }
}

// pt_count[i] = #selected models containing point i ; model_count[b] += 1 for each hit.
// If pou != nullptr also counts rows: number of maximal runs of equal pou among the hits (ascending model).
template <int XD>
__global__ void __launch_bounds__(TT)
inside_count_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ sub_static, int ss,
                    const int32_t* __restrict__ models, int n_models, const int32_t* __restrict__ pou,
                    const int32_t* __restrict__ pos, int32_t* __restrict__ pt_count, int32_t* __restrict__ pt_rows, int32_t* __restrict__ model_count) {
    __shared__ float s_box[2 * XD * BOX_CHUNK];
    __shared__ int32_t s_pou[BOX_CHUNK];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float xx[XD];
#pragma unroll
    for (int d = 0; d < XD; ++d) xx[d] = i < n ? x[i * XD + d] : 0.0f;
    int cnt = 0, rows = 0, prev = -1;
    for (int b0 = 0; b0 < n_models; b0 += BOX_CHUNK) {
        int nb = min(BOX_CHUNK, n_models - b0);
        __syncthreads();
        stage_boxes<XD>(s_box, pou ? s_pou : nullptr, sub_static, models, pou, pos, b0, nb, ss);
        __syncthreads();
        if (i < n) {
            for (int b = 0; b < nb; ++b) {
                if (inside_box<XD>(s_box, b, xx)) {
                    ++cnt;
                    if (model_count) atomicAdd(model_count + b0 + b, 1);
                    if (pou) {
                        int pv = s_pou[b];
                        if (pv != prev) { ++rows; prev = pv; }
                    }
                }
            }
        }
    }
    if (i < n) {
        pt_count[i] = cnt;
        if (pt_rows) pt_rows[i] = rows;
    }
}

// Second pass: write the reference-order arrays from the per-point offsets.
template <int XD>
__global__ void __launch_bounds__(TT)
takes_fill_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ sub_static, int ss, int m,
                  const int32_t* __restrict__ pos_of_model, const int32_t* __restrict__ pou,
                  const int32_t* __restrict__ pt_off, const int32_t* __restrict__ pt_row_off,
                  int32_t* __restrict__ m_take, int32_t* __restrict__ n_take, int32_t* __restrict__ p_take,
                  int32_t* __restrict__ np_take, int32_t* __restrict__ row_off) {
    __shared__ float s_box[2 * XD * BOX_CHUNK];
    __shared__ int32_t s_pou[BOX_CHUNK];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float xx[XD];
#pragma unroll
    for (int d = 0; d < XD; ++d) xx[d] = i < n ? x[i * XD + d] : 0.0f;
    int j = i < n ? pt_off[i] : 0;
    int r = i < n ? pt_row_off[i] - 1 : 0;
    int prev = -1;
    for (int b0 = 0; b0 < m; b0 += BOX_CHUNK) {
        int nb = min(BOX_CHUNK, m - b0);
        __syncthreads();
        stage_boxes<XD>(s_box, s_pou, sub_static, nullptr, pou, pos_of_model, b0, nb, ss);
        __syncthreads();
        if (i < n) {
            for (int b = 0; b < nb; ++b) {
                if (inside_box<XD>(s_box, b, xx)) {
                    int pv = s_pou[b];
                    if (pv != prev) {
                        ++r;
                        prev = pv;
                        np_take[r] = (int32_t)i;
                        row_off[r] = j;
                    }
                    n_take[j] = (int32_t)i;
                    m_take[j] = pos_of_model[b0 + b];
                    p_take[j] = r;
                    ++j;
                }
            }
        }
    }
}

__global__ void iota_kernel(int32_t* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

__global__ void hist_kernel(const int32_t* __restrict__ keys, int64_t n, int32_t* __restrict__ hist) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(hist + keys[i], 1);
}

// sorted_idx[i] = reference-order pair index sitting at subdomain-sorted slot i
__global__ void sorted_view_kernel(const int32_t* __restrict__ sorted_idx, int64_t s, const int32_t* __restrict__ n_take,
                                   const int32_t* __restrict__ m_take, const int32_t* __restrict__ p_take,
                                   int32_t* __restrict__ spair_point, int32_t* __restrict__ spair_row,
                                   int32_t* __restrict__ spair_sub, int32_t* __restrict__ pos) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s) return;
    int j = sorted_idx[i];
    spair_point[i] = n_take[j];
    spair_row[i] = p_take[j];
    spair_sub[i] = m_take[j];
    pos[j] = (int32_t)i;
}

__global__ void set_i32_kernel(int32_t* p, int32_t v) { *p = v; }

inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

int exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, int64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    void* tmp = nullptr;
    size_t bytes = 0;
    FBP_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_in, d_out, (int)n, st));
    FBP_CHECK_CUDA(cudaMalloc(&tmp, bytes ? bytes : 1));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, d_in, d_out, (int)n, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(tmp);
    FBP_CHECK_CUDA(e);
    FBP_CHECK_CUDA(e2);
    return 0;
}

template <int XD>
int launch_count(const float* d_x, int64_t n, const float* d_sub_static, const int32_t* d_models, int n_models,
                 const int32_t* d_pou, const int32_t* d_pos, int32_t* d_pt_count, int32_t* d_pt_rows,
                 int32_t* d_model_count, cudaStream_t st) {
    inside_count_kernel<XD><<<nblk(n, TT), TT, 0, st>>>(d_x, n, d_sub_static, 2 * XD + 3, d_models, n_models, d_pou, d_pos,
                                                        d_pt_count, d_pt_rows, d_model_count);
    FBP_LAUNCH_CHECK();
    return 0;
}

}  // namespace

struct fbp_takes_builder {
    const float* d_x;
    int64_t n;
    int xd, m, m_all;
    const float* d_sub_static;
    const int32_t* d_pos_of_model;
    const int32_t* d_pou_of_model;
    int32_t* d_pt_off = nullptr;      // [n+1] exclusive scan of pair counts
    int32_t* d_pt_row_off = nullptr;  // [n+1] exclusive scan of row counts
    int64_t s = 0, q = 0;
};

extern "C" {

int fbp_inside_count(const float* d_x, int64_t n, int32_t xd, const float* d_sub_static, int32_t m,
                     const int32_t* d_models, int32_t n_models, int32_t* d_pt_count, int32_t* d_model_count,
                     void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FBP_REQUIRE(xd >= 1 && xd <= FBP_MAX_XD, "fbp_inside_count: xd=%d unsupported (1..%d)", xd, FBP_MAX_XD);
    FBP_REQUIRE(n < (1ll << 31), "fbp_inside_count: n too large for int32 indices");
    if (!d_models) n_models = m;
    if (d_model_count && n_models > 0)
        FBP_CHECK_CUDA(cudaMemsetAsync(d_model_count, 0, sizeof(int32_t) * n_models, st));
    if (n == 0) return 0;
    switch (xd) {
        case 1: return launch_count<1>(d_x, n, d_sub_static, d_models, n_models, nullptr, nullptr, d_pt_count, nullptr, d_model_count, st);
        case 2: return launch_count<2>(d_x, n, d_sub_static, d_models, n_models, nullptr, nullptr, d_pt_count, nullptr, d_model_count, st);
        default: return launch_count<3>(d_x, n, d_sub_static, d_models, n_models, nullptr, nullptr, d_pt_count, nullptr, d_model_count, st);
    }
}

int fbp_nonzero_i32(const int32_t* d_count, int64_t n, int32_t* d_out, int64_t* n_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    *n_out = 0;
    if (n == 0) return 0;
    FBP_REQUIRE(n < (1ll << 31), "fbp_nonzero_i32: n too large");
    int32_t* d_iota = nullptr;
    int32_t* d_num = nullptr;
    void* tmp = nullptr;
    size_t bytes = 0;
    FBP_CHECK_CUDA(cudaMalloc(&d_iota, sizeof(int32_t) * n));
    FBP_CHECK_CUDA(cudaMalloc(&d_num, sizeof(int32_t)));
    iota_kernel<<<nblk(n, TT), TT, 0, st>>>(d_iota, n);
    cudaError_t e = cub::DeviceSelect::Flagged(nullptr, bytes, d_iota, d_count, d_out, d_num, (int)n, st);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, bytes ? bytes : 1);
    if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(tmp, bytes, d_iota, d_count, d_out, d_num, (int)n, st);
    int32_t num = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&num, d_num, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    cudaFree(d_iota);
    cudaFree(d_num);
    FBP_CHECK_CUDA(e);
    *n_out = num;
    return 0;
}

int fbp_takes_begin(fbp_takes_builder** out, const float* d_x, int64_t n, int32_t xd, const float* d_sub_static,
                    int32_t m, const int32_t* d_pos_of_model, const int32_t* d_pou_of_model, int32_t m_all,
                    void* stream, int64_t* s_out, int64_t* q_out) {
    cudaStream_t st = (cudaStream_t)stream;
    FBP_REQUIRE(out && s_out && q_out, "fbp_takes_begin: null output");
    FBP_REQUIRE(xd >= 1 && xd <= FBP_MAX_XD, "fbp_takes_begin: xd=%d unsupported (1..%d)", xd, FBP_MAX_XD);
    FBP_REQUIRE(n < (1ll << 31) - 1, "fbp_takes_begin: n too large for int32 indices");
    fbp_takes_builder* b = new fbp_takes_builder();
    b->d_x = d_x; b->n = n; b->xd = xd; b->m = m; b->m_all = m_all;
    b->d_sub_static = d_sub_static; b->d_pos_of_model = d_pos_of_model; b->d_pou_of_model = d_pou_of_model;
    *out = b;
    *s_out = 0; *q_out = 0;
    FBP_CHECK_CUDA(cudaMalloc(&b->d_pt_off, sizeof(int32_t) * (n + 1)));
    FBP_CHECK_CUDA(cudaMalloc(&b->d_pt_row_off, sizeof(int32_t) * (n + 1)));
    if (n == 0) {
        FBP_CHECK_CUDA(cudaMemsetAsync(b->d_pt_off, 0, sizeof(int32_t), st));
        FBP_CHECK_CUDA(cudaMemsetAsync(b->d_pt_row_off, 0, sizeof(int32_t), st));
        FBP_CHECK_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    int32_t *d_cnt = nullptr, *d_rows = nullptr;
    FBP_CHECK_CUDA(cudaMalloc(&d_cnt, sizeof(int32_t) * (n + 1)));
    FBP_CHECK_CUDA(cudaMalloc(&d_rows, sizeof(int32_t) * (n + 1)));
    FBP_CHECK_CUDA(cudaMemsetAsync(d_cnt + n, 0, sizeof(int32_t), st));
    FBP_CHECK_CUDA(cudaMemsetAsync(d_rows + n, 0, sizeof(int32_t), st));
    int rc;
    switch (xd) {
        case 1: rc = launch_count<1>(d_x, n, d_sub_static, nullptr, m, d_pou_of_model, d_pos_of_model, d_cnt, d_rows, nullptr, st); break;
        case 2: rc = launch_count<2>(d_x, n, d_sub_static, nullptr, m, d_pou_of_model, d_pos_of_model, d_cnt, d_rows, nullptr, st); break;
        default: rc = launch_count<3>(d_x, n, d_sub_static, nullptr, m, d_pou_of_model, d_pos_of_model, d_cnt, d_rows, nullptr, st); break;
    }
    if (rc == 0) rc = exclusive_scan_i32(d_cnt, b->d_pt_off, n + 1, st);
    if (rc == 0) rc = exclusive_scan_i32(d_rows, b->d_pt_row_off, n + 1, st);
    int32_t tot[2] = {0, 0};
    if (rc == 0) {
        cudaError_t e = cudaMemcpyAsync(&tot[0], b->d_pt_off + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&tot[1], b->d_pt_row_off + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { fbp_set_error("fbp_takes_begin: %s", cudaGetErrorString(e)); rc = 1; }
    }
    cudaFree(d_cnt);
    cudaFree(d_rows);
    if (rc) return rc;
    FBP_REQUIRE(tot[0] >= 0 && tot[1] >= 0, "fbp_takes_begin: pair count overflows int32");
    b->s = tot[0];
    b->q = tot[1];
    *s_out = b->s;
    *q_out = b->q;
    return 0;
}

int fbp_takes_emit(fbp_takes_builder* b, int32_t* d_m_take, int32_t* d_n_take, int32_t* d_p_take, int32_t* d_np_take,
                   int32_t* d_row_off, int32_t* d_pt_row_off, int32_t* d_sub_off, int32_t* d_spair_point,
                   int32_t* d_spair_row, int32_t* d_spair_sub, int32_t* d_pos, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FBP_REQUIRE(b, "fbp_takes_emit: null builder");
    const int64_t n = b->n, s = b->s, q = b->q;
    FBP_CHECK_CUDA(cudaMemcpyAsync(d_pt_row_off, b->d_pt_row_off, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToDevice, st));
    set_i32_kernel<<<1, 1, 0, st>>>(d_row_off + q, (int32_t)s);
    FBP_LAUNCH_CHECK();
    FBP_CHECK_CUDA(cudaMemsetAsync(d_sub_off, 0, sizeof(int32_t) * (b->m_all + 1), st));
    if (s == 0) {
        FBP_CHECK_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    switch (b->xd) {
        case 1: takes_fill_kernel<1><<<nblk(n, TT), TT, 0, st>>>(b->d_x, n, b->d_sub_static, 5, b->m, b->d_pos_of_model, b->d_pou_of_model, b->d_pt_off, b->d_pt_row_off, d_m_take, d_n_take, d_p_take, d_np_take, d_row_off); break;
        case 2: takes_fill_kernel<2><<<nblk(n, TT), TT, 0, st>>>(b->d_x, n, b->d_sub_static, 7, b->m, b->d_pos_of_model, b->d_pou_of_model, b->d_pt_off, b->d_pt_row_off, d_m_take, d_n_take, d_p_take, d_np_take, d_row_off); break;
        default: takes_fill_kernel<3><<<nblk(n, TT), TT, 0, st>>>(b->d_x, n, b->d_sub_static, 9, b->m, b->d_pos_of_model, b->d_pou_of_model, b->d_pt_off, b->d_pt_row_off, d_m_take, d_n_take, d_p_take, d_np_take, d_row_off); break;
    }
    FBP_LAUNCH_CHECK();

    // subdomain-sorted view: stable sort of pair indices by subdomain position (keeps points ascending)
    int32_t *d_hist = nullptr, *d_iota = nullptr, *d_keys_out = nullptr, *d_vals_out = nullptr;
    void* tmp = nullptr;
    size_t bytes = 0;
    int rc = 0;
    cudaError_t e = cudaMalloc(&d_hist, sizeof(int32_t) * (b->m_all + 1));
    if (e == cudaSuccess) e = cudaMalloc(&d_iota, sizeof(int32_t) * s);
    if (e == cudaSuccess) e = cudaMalloc(&d_keys_out, sizeof(int32_t) * s);
    if (e == cudaSuccess) e = cudaMalloc(&d_vals_out, sizeof(int32_t) * s);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_hist, 0, sizeof(int32_t) * (b->m_all + 1), st);
    if (e == cudaSuccess) {
        hist_kernel<<<nblk(s, TT), TT, 0, st>>>(d_m_take, s, d_hist);
        iota_kernel<<<nblk(s, TT), TT, 0, st>>>(d_iota, s);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        rc = exclusive_scan_i32(d_hist, d_sub_off, b->m_all + 1, st);
        int end_bit = 1;
        while ((1ll << end_bit) < (long long)b->m_all + 1 && end_bit < 31) ++end_bit;
        if (rc == 0) e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_m_take, d_keys_out, d_iota, d_vals_out, (int)s, 0, end_bit, st);
        if (rc == 0 && e == cudaSuccess) e = cudaMalloc(&tmp, bytes ? bytes : 1);
        if (rc == 0 && e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, bytes, d_m_take, d_keys_out, d_iota, d_vals_out, (int)s, 0, end_bit, st);
        if (rc == 0 && e == cudaSuccess) {
            sorted_view_kernel<<<nblk(s, TT), TT, 0, st>>>(d_vals_out, s, d_n_take, d_m_take, d_p_take, d_spair_point,
                                                          d_spair_row, d_spair_sub, d_pos);
            e = cudaGetLastError();
        }
    }
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(tmp); cudaFree(d_hist); cudaFree(d_iota); cudaFree(d_keys_out); cudaFree(d_vals_out);
    if (rc) return rc;
    FBP_CHECK_CUDA(e);
    FBP_CHECK_CUDA(e2);
    return 0;
}

int fbp_takes_destroy(fbp_takes_builder* b) {
    if (!b) return 0;
    cudaFree(b->d_pt_off);
    cudaFree(b->d_pt_row_off);
    delete b;
    return 0;
}

}  // extern "C"
