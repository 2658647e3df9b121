// Gather roofline micro-benchmark: how fast can a B200 gather whole fp32
// feature rows when there is NO sparse-matrix bookkeeping at all?
//
// The kernel mirrors spmm_rows_kernel's access pattern (one warp per output row, `deg` gathered
// rows of `F` floats each, 8 gathers in flight per warp, float4 lanes, accumulate, one streamed
// output row) but takes its column ids from a hash instead of a CSR, so what is measured is the
// memory system: L1 / L2->SM fabric / DRAM under row gathers of the products shape.
//
// Column distribution (per entry):  with probability p_uni the column is uniform over all rows
// (always a DRAM miss once H >> L2); otherwise |col - row| is log-uniform in [1, dmax] — the
// distance law of the synthetic products graph (gnn-tf_b200/synthetic.py:powerlaw_edges), so
// dmax = N/2 reproduces its locality, a small dmax gives an all-L1/L2-hit gather and p_uni = 1 an
// all-miss gather.
//
// Output: one JSON object per configuration on stdout.  Build: make -C scripts/microbench.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                            \
    do {                                                                                 \
        cudaError_t e_ = (x);                                                            \
        if (e_ != cudaSuccess) {                                                         \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                     \
        }                                                                                \
    } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// column of entry e of row `row`
__device__ __forceinline__ int pick_col(int row, int e, int n, float log2_dmax, uint32_t uni_thresh, uint32_t seed) {
    const uint32_t h = mix((uint32_t)row * 0x9e3779b9u + (uint32_t)e * 0x85ebca6bu + seed);
    const uint32_t h2 = mix(h + 0x68bc21ebu);
    if (h2 < uni_thresh) return (int)(((uint64_t)h * (uint64_t)n) >> 32);
    const float u = (float)(h >> 8) * (1.0f / 16777216.0f);
    int d = (int)exp2f(u * log2_dmax);
    d = max(1, d);
    int c = row + ((h2 & 1u) ? d : -d);
    if (c < 0) c += n;
    if (c >= n) c -= n;
    return c;
}

// LANES active lanes per warp own float4 slot `lane` of every gathered row (F = 4*LANES).
template <int UNROLL>
__global__ void __launch_bounds__(256, 5)
gather_ldg_kernel(const float* __restrict__ H, float* __restrict__ out, int n, int ld, int lanes, int deg,
                  float log2_dmax, uint32_t uni_thresh, uint32_t seed, int rows_per_cta, int write_out) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool on = lane < lanes;
    const float* Hl = H + (on ? lane * 4 : 0);
    const uint32_t pitch = (uint32_t)ld * 4u;
    for (int it = 0; it < rows_per_cta; it += 8) {
        const int row = blockIdx.x * rows_per_cta + it + warp;
        if (row >= n) break;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = 0; base < deg; base += 32) {
            const int mycol = pick_col(row, base + lane, n, log2_dmax, uni_thresh, seed);
            const int cnt = min(32, deg - base);
            for (int j = 0; j < cnt; j += UNROLL) {
                float4 x[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const int c = __shfl_sync(0xffffffffu, mycol, (j + u) & 31);
                    const float* p;
                    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(c), "r"(pitch), "l"(Hl));
                    if (on) x[u] = __ldg(reinterpret_cast<const float4*>(p));
                    else x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const float w = (j + u < cnt) ? 0.02f : 0.0f;
                    acc.x = fmaf(w, x[u].x, acc.x); acc.y = fmaf(w, x[u].y, acc.y);
                    acc.z = fmaf(w, x[u].z, acc.z); acc.w = fmaf(w, x[u].w, acc.w);
                }
            }
        }
        if (on) {
            if (write_out) {
                float* o = out + (size_t)row * ld + lane * 4;
                asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(acc.x), "f"(acc.y),
                             "f"(acc.z), "f"(acc.w)
                             : "memory");
            } else if (acc.x == 123.456f) {
                out[0] = acc.y;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Bisecting the gap between the bare gather above and the real SpMM kernel: the same loop fed from
// real arrays.  cols/vals: [n*deg] (fixed degree) or CSR with row_ptr (variable degree);
// WITH_H0: the epilogue also streams a teleport row in.
// ------------------------------------------------------------------------------------------------
__global__ void fill_cols_kernel(int* cols, float* vals, const int* row_ptr, int n, int deg, float log2_dmax, uint32_t seed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)n * 64; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i >> 6), e = (int)(i & 63);
        const int s = row_ptr ? row_ptr[row] : row * deg;
        const int d = row_ptr ? row_ptr[row + 1] - s : deg;
        if (e < d) {
            cols[s + e] = pick_col(row, e, n, log2_dmax, 0, seed);
            vals[s + e] = 0.02f;
        }
    }
}

template <bool WITH_H0, bool VAR_DEG>
__global__ void __launch_bounds__(256, 5)
gather_csr_kernel(const float* __restrict__ H, const float* __restrict__ H0, float* __restrict__ out,
                  const int* __restrict__ row_ptr, const int* __restrict__ cols, const float* __restrict__ vals,
                  int n, int ld, int lanes, int deg_fixed, int rows_per_cta) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool on = lane < lanes;
    const float* Hl = H + (on ? lane * 4 : 0);
    const uint32_t pitch = (uint32_t)ld * 4u;
    for (int it = 0; it < rows_per_cta; it += 8) {
        const int row = blockIdx.x * rows_per_cta + it + warp;
        if (row >= n) break;
        int start = row * deg_fixed, deg = deg_fixed;
        if (VAR_DEG) {
            start = __ldg(row_ptr + row);
            deg = __ldg(row_ptr + row + 1) - start;
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = 0; base < deg; base += 32) {
            int mycol = 0;
            float myval = 0.f;
            if (base + lane < deg) {
                asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(mycol) : "l"(cols + start + base + lane));
                asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(myval) : "l"(vals + start + base + lane));
            }
            const int cnt = min(32, deg - base);
            for (int j = 0; j < cnt; j += 8) {
                float4 x[8];
                float w[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = __shfl_sync(0xffffffffu, mycol, (j + u) & 31);
                    w[u] = __shfl_sync(0xffffffffu, myval, (j + u) & 31);
                    const float* p;
                    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(c), "r"(pitch), "l"(Hl));
                    if (on) x[u] = __ldg(reinterpret_cast<const float4*>(p));
                    else x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    acc.x = fmaf(w[u], x[u].x, acc.x); acc.y = fmaf(w[u], x[u].y, acc.y);
                    acc.z = fmaf(w[u], x[u].z, acc.z); acc.w = fmaf(w[u], x[u].w, acc.w);
                }
            }
        }
        if (on) {
            if (WITH_H0) {
                float4 h;
                const float* hp = H0 + (size_t)row * ld + lane * 4;
                asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(h.x), "=f"(h.y), "=f"(h.z), "=f"(h.w) : "l"(hp));
                acc.x = acc.x * 0.9f + h.x * 0.1f; acc.y = acc.y * 0.9f + h.y * 0.1f;
                acc.z = acc.z * 0.9f + h.z * 0.1f; acc.w = acc.w * 0.9f + h.w * 0.1f;
            }
            float* o = out + (size_t)row * ld + lane * 4;
            asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(acc.x), "f"(acc.y),
                         "f"(acc.z), "f"(acc.w) : "memory");
        }
    }
}

struct Cfg {
    const char* name;
    int n, F, ld, deg;
    double dmax;   // <= 0: N/2
    double p_uni;
    int write_out;
    int rows_per_cta;
};

int main(int argc, char** argv) {
    int reps = 5;
    const char* filter = nullptr;  // run only configurations whose name contains this substring
    if (argc > 1) reps = atoi(argv[1]);
    if (argc > 2) filter = argv[2];
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    fprintf(stderr, "device %s, %d SMs, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);

    const int NP = 2449029;  // products rows
    std::vector<Cfg> cfgs = {
        // ---- ceilings
        {"l2_hit_small_H_F100", 80000, 100, 100, 50, 0, 1.0, 1, 64},        // H = 32 MB: every gather an L2 hit
        {"l2_hit_small_H_F128", 80000, 128, 128, 50, 0, 1.0, 1, 64},
        {"l1_hit_band16_F100", NP, 100, 100, 50, 16, 0.0, 1, 64},            // +-16 rows: L1-resident band
        {"l1l2_band256_F100", NP, 100, 100, 50, 256, 0.0, 1, 64},
        {"l2_band4k_F100", NP, 100, 100, 50, 4096, 0.0, 1, 64},
        {"l2_band32k_F100", NP, 100, 100, 50, 32768, 0.0, 1, 64},
        {"dram_miss_uniform_F100", NP, 100, 100, 50, 0, 1.0, 1, 64},         // H = 980 MB uniform: ~all DRAM misses
        {"dram_miss_uniform_F100_ld128", NP, 100, 128, 50, 0, 1.0, 1, 64},   // aligned 512-byte pitch
        {"dram_miss_uniform_F128", NP, 128, 128, 50, 0, 1.0, 1, 64},
        {"dram_miss_uniform_F100_nowrite", NP, 100, 100, 50, 0, 1.0, 0, 64},
        // ---- the products distance law (log-uniform up to N/2) and mixtures
        {"products_law_F100", NP, 100, 100, 50, -1, 0.0, 1, 64},
        {"products_law_F100_rpc8", NP, 100, 100, 50, -1, 0.0, 1, 8},
        {"products_law_F100_rpc512", NP, 100, 100, 50, -1, 0.0, 1, 512},
        {"products_law_F128", NP, 128, 128, 50, -1, 0.0, 1, 64},
        {"band32k_plus_10pct_uniform", NP, 100, 100, 50, 32768, 0.10, 1, 64},
        {"band32k_plus_20pct_uniform", NP, 100, 100, 50, 32768, 0.20, 1, 64},
        {"band32k_plus_35pct_uniform", NP, 100, 100, 50, 32768, 0.35, 1, 64},
        // ---- the arxiv shape (everything L2-resident)
        {"arxiv_law_F128", 169343, 128, 128, 14, -1, 0.0, 1, 8},
        {"arxiv_law_F40", 169343, 40, 40, 14, -1, 0.0, 1, 8},
    };

    size_t max_elems = 0;
    for (auto& c : cfgs) max_elems = std::max(max_elems, (size_t)c.n * c.ld);
    float *H, *out;
    CK(cudaMalloc(&H, max_elems * sizeof(float)));
    CK(cudaMalloc(&out, max_elems * sizeof(float)));
    CK(cudaMemset(H, 0, max_elems * sizeof(float)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    for (auto& c : cfgs) {
        if (filter && !strstr(c.name, filter)) continue;
        const double dmax = c.dmax > 0 ? c.dmax : c.n / 2.0;
        const float log2_dmax = (float)std::log2(dmax);
        const uint32_t thr = c.p_uni >= 1.0 ? 0xffffffffu : (uint32_t)(c.p_uni * 4294967296.0);
        const int lanes = c.F / 4;
        const int grid = (c.n + c.rows_per_cta - 1) / c.rows_per_cta;
        std::vector<float> ms;
        for (int r = 0; r < reps + 2; ++r) {
            CK(cudaEventRecord(e0));
            gather_ldg_kernel<8><<<grid, 256>>>(H, out, c.n, c.ld, lanes, c.deg, log2_dmax, thr, 1234u + r,
                                                c.rows_per_cta, c.write_out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float t;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (r >= 2) ms.push_back(t);
        }
        std::sort(ms.begin(), ms.end());
        const double med = ms[ms.size() / 2];
        const double gathers = (double)c.n * c.deg;
        const double gbytes = gathers * c.F * 4.0;
        printf("{\"bench\": \"gather_ldg\", \"name\": \"%s\", \"n\": %d, \"F\": %d, \"ld\": %d, \"deg\": %d, "
               "\"dmax\": %.0f, \"p_uniform\": %.2f, \"write_out\": %d, \"rows_per_cta\": %d, \"ms\": %.4f, \"ms_min\": %.4f, "
               "\"gather_rows_per_s\": %.4g, \"gather_GBps\": %.1f, \"H_MB\": %.1f}\n",
               c.name, c.n, c.F, c.ld, c.deg, dmax, c.p_uni, c.write_out, c.rows_per_cta, med, ms[0],
               gathers / (med * 1e-3), gbytes / (med * 1e-3) / 1e9, (double)c.n * c.ld * 4 / 1e6);
        fflush(stdout);
    }

    // ---- bisect: the same gather fed from real arrays (products law, F=100) ----
    if (!filter || strstr("bisect_csr", filter)) {
        const int n = NP, deg = 50, F = 100, ld = 100;
        const float l2d = (float)std::log2(n / 2.0);
        int *cols, *row_ptr;
        float *vals, *H0;
        CK(cudaMalloc(&cols, (size_t)n * 64 * 4));
        CK(cudaMalloc(&vals, (size_t)n * 64 * 4));
        CK(cudaMalloc(&row_ptr, (size_t)(n + 1) * 4));
        CK(cudaMalloc(&H0, (size_t)n * ld * 4));
        CK(cudaMemset(H0, 0, (size_t)n * ld * 4));
        // variable degrees: 25 + binomial-ish spread in [25, 75], mean 50
        std::vector<int> rp(n + 1);
        rp[0] = 0;
        uint32_t st = 12345u;
        for (int i = 0; i < n; ++i) {
            int d = 25;
            for (int k = 0; k < 5; ++k) { st = st * 1664525u + 1013904223u; d += (st >> 24) % 11; }
            rp[i + 1] = rp[i] + d;
        }
        CK(cudaMemcpy(row_ptr, rp.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice));
        for (int variant = 0; variant < 4; ++variant) {
            const bool var_deg = variant == 3;
            fill_cols_kernel<<<148 * 8, 256>>>(cols, vals, var_deg ? row_ptr : nullptr, n, deg, l2d, 77u);
            CK(cudaDeviceSynchronize());
            std::vector<float> ms;
            const int grid = (n + 63) / 64;
            for (int r = 0; r < reps + 2; ++r) {
                CK(cudaEventRecord(e0));
                if (variant == 0) gather_ldg_kernel<8><<<grid, 256>>>(H, out, n, ld, F / 4, deg, l2d, 0, 77u, 64, 1);
                else if (variant == 1) gather_csr_kernel<false, false><<<grid, 256>>>(H, H0, out, row_ptr, cols, vals, n, ld, F / 4, deg, 64);
                else if (variant == 2) gather_csr_kernel<true, false><<<grid, 256>>>(H, H0, out, row_ptr, cols, vals, n, ld, F / 4, deg, 64);
                else gather_csr_kernel<true, true><<<grid, 256>>>(H, H0, out, row_ptr, cols, vals, n, ld, F / 4, deg, 64);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float t;
                CK(cudaEventElapsedTime(&t, e0, e1));
                if (r >= 2) ms.push_back(t);
            }
            std::sort(ms.begin(), ms.end());
            const char* vn[] = {"hashed columns (bare gather)", "+ (col,val) read from arrays", "+ teleport row read", "+ variable degree 25..75 (row_ptr)"};
            const double entries = var_deg ? (double)rp[n] : (double)n * deg;
            printf("{\"bench\": \"bisect_csr\", \"variant\": \"%s\", \"entries\": %.0f, \"ms\": %.4f, \"ns_per_1k_entries\": %.3f}\n", vn[variant],
                   entries, ms[ms.size() / 2], ms[ms.size() / 2] * 1e6 / entries * 1e3 / 1e3);
            fflush(stdout);
        }
    }
    return 0;
}
