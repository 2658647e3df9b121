// CSR SpMM for sm_100a: row-split main kernel + split long rows (nnz-chunked, fixed-order reduce).
//
// Replaces tf.sparse.sparse_dense_matmul at gnntf/core/gnn/architectures/filter.py:19 and
// gcn.py:24,48,88,104,131, fused with the teleport axpy / dropout / activation of
// filter.py:21-22.  The path is bandwidth-bound (AI < 7 flop/B): no tensor cores.
//
// Lane mapping.  A feature row of F floats is cut into VEC-wide slots (VEC = 4 -> LDG.128 when
// every leading dimension and base pointer allows it, else scalar).  GROUP lanes (a power of two
// <= 32) own one sparse row; lane `gl` of the group owns slots gl, gl+GROUP, ... (NSLOT of them),
// so the lanes of a group read one contiguous GROUP*VEC*4-byte span of each gathered row.  A warp
// therefore runs 32/GROUP sparse rows at once.  The group's lanes fetch GROUP (col,val) pairs with
// one coalesced streaming load each and broadcast them through a small per-warp shared-memory
// buffer; the gathers of UNROLL consecutive entries are issued back to back before the FMAs
// (memory-level parallelism).  Rows wider than 128 floats run as 128-float tiles over grid.y.
//
// Accumulation order inside a row is CSR slot order = COO storage order (the builder's sort is
// stable), i.e. the order TF's CPU kernel uses.  Rows longer than A->long_threshold are skipped
// by the row warps and run as fixed-size pieces on other warps of the same grid (process_piece),
// whose partial sums are combined in piece order by spmm_long_reduce_kernel: deterministic, no
// float atomics.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include <cooperative_groups.h>

#include "push.cuh"
#include "spmm.cuh"

namespace gnntf {

// `h` is the teleport row slice H0[row, f..f+VEC) when the caller already loaded it (the row kernel
// issues that load before the row's gathers so its latency is hidden); it is ignored when e.H0 == NULL.
template <int VEC>
__device__ __forceinline__ void apply_epilogue_h(const Epilogue& e, int64_t row, int f, Vec<VEC> acc, Vec<VEC> h) {
    Vec<VEC> out;
#pragma unroll
    for (int i = 0; i < VEC; ++i) out.v[i] = acc.v[i] * e.s;
    if (e.H0 != nullptr) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) out.v[i] = __fadd_rn(out.v[i], __fmul_rn(h.v[i], e.t));
    }
    if (e.keep != nullptr) {
        const uint8_t* k = e.keep + row * (int64_t)e.F + f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) out.v[i] = k[i] ? out.v[i] * e.p_scale : 0.0f;
    }
    if (e.act == GNNTF_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) out.v[i] = fmaxf(out.v[i], 0.0f);
    }
    if (e.C != nullptr) out.store(e.C + row * e.ldc + f);
    if (e.ACC != nullptr) {
        float* a = e.ACC + row * e.ldacc + f;
        Vec<VEC> b;
#pragma unroll
        for (int i = 0; i < VEC; ++i) b.v[i] = 0.0f;
        if (e.u != 0.0f) b = Vec<VEC>::stream(e.B + row * e.ldb + f);  // the dense operand's own row
        Vec<VEC> r;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float prev = e.acc_init ? 0.0f : a[i];
            r.v[i] = prev + e.u * b.v[i] + e.w * out.v[i];
        }
        r.store(a);
    }
}

template <int VEC>
__device__ __forceinline__ void apply_epilogue(const Epilogue& e, int64_t row, int f, Vec<VEC> acc) {
    Vec<VEC> h;
#pragma unroll
    for (int i = 0; i < VEC; ++i) h.v[i] = 0.0f;
    if (e.H0 != nullptr) h = Vec<VEC>::stream(e.H0 + row * e.ldh + f);
    apply_epilogue_h<VEC>(e, row, f, acc, h);
}

// L2 eviction-policy helpers: data touched once per step (CSR arrays, the teleport term, the
// output) is marked evict-first so it does not displace gathered feature rows from L2.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int ld_once(const int* p, uint64_t pol) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_once(const float* p, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ld_once4(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_once4(float* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}


// ---------------------------------------------------------------------------------------------
// Long rows, phase 1: one warp per 256-entry piece.  The 32/GROUP lane groups take the piece's
// entries round-robin and are combined with xor-shuffles in a fixed pattern; the result goes to
// partials[piece, ldp] (ldp = round_up(F,4)).  Runs inside the row kernel's grid (the first
// `piece_ctas` CTAs), so the pieces overlap the ordinary rows instead of being a second launch.
// ---------------------------------------------------------------------------------------------
struct PieceArgs {
    const int* chunk_row;
    const int* chunk_begin;
    int n_chunks;
    int chunk;
    float* partials;
    int ldp;
};

template <int VEC, int NSLOT, int GROUP, int UNROLL>
__device__ __forceinline__ void process_piece(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                              const float* __restrict__ val, const float* __restrict__ B,
                                              int64_t ldb, const int* __restrict__ chunk_row,
                                              const int* __restrict__ chunk_begin, int piece, int chunk,
                                              float* __restrict__ partials, int ldp, int F, int f_base) {
    constexpr int NGROUPS = 32 / GROUP;
    const int lane = threadIdx.x & 31;
    const int g = lane / GROUP;
    const int gl = lane % GROUP;
    const int row = __ldg(chunk_row + piece);
    const int begin = __ldg(chunk_begin + piece);
    const int end = min(begin + chunk, __ldg(row_ptr + row + 1));

    int fo[NSLOT];
    bool fok[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        fo[s] = f_base + (s * GROUP + gl) * VEC;
        fok[s] = fo[s] < F;
    }
    Vec<VEC> acc[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[s].v[i] = 0.0f;

    // 32 entries per outer step: lane l holds entry base+l; group g consumes entries with
    // (index % NGROUPS) == g.
    for (int base = begin; base < end; base += 32) {
        int c = 0;
        float v = 0.0f;
        if (base + lane < end) {
            c = ld_stream(col_idx + base + lane);
            v = ld_stream(val + base + lane);
        }
        const int cnt = min(32, end - base);
        for (int j = 0; j < cnt; j += NGROUPS * UNROLL) {
            int cj[UNROLL];
            float vj[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int src = j + u * NGROUPS + g;
                cj[u] = __shfl_sync(0xffffffffu, c, src & 31);
                vj[u] = __shfl_sync(0xffffffffu, v, src & 31);
                if (src >= cnt) vj[u] = 0.0f, cj[u] = -1;
            }
            Vec<VEC> x[UNROLL][NSLOT];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float* src = B + (int64_t)cj[u] * ldb;
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    if (cj[u] >= 0 && fok[s]) {
                        x[u][s] = Vec<VEC>::gather(src + fo[s]);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) x[u][s].v[i] = 0.0f;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[s].v[i] = fmaf(vj[u], x[u][s].v[i], acc[s].v[i]);
        }
    }
    if (NGROUPS > 1) {
#pragma unroll
        for (int o = GROUP; o < 32; o <<= 1)
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[s].v[i] += __shfl_xor_sync(0xffffffffu, acc[s].v[i], o);
    }
    if (g == 0) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
            if (fok[s]) acc[s].store(partials + (int64_t)piece * ldp + fo[s]);
    }
}

// ---------------------------------------------------------------------------------------------
// Main kernel: GROUP lanes per sparse row; rows with deg > long_threshold are left to the piece
// path.  A CTA walks blocks_per_cta consecutive row blocks (fewer, longer-lived CTAs: the per-row
// version launched 306 k CTAs on the products shape and ran at 47 % achieved occupancy);
// grid.y walks feature tiles of GROUP*NSLOT*VEC floats.
// A gather address costs one mad.wide.u32 (the first profile showed 28 executed instructions per
// entry, most of them 64-bit address arithmetic, per-entry predicates and register zeroing); every
// batch of UNROLL entries runs predicate-free (short rows are padded, see the loop).
// ---------------------------------------------------------------------------------------------

// Address of a gathered row piece in ONE instruction: column ids are non-negative int32 and the
// row pitch in bytes fits 32 bits, so base + c*pitch is exactly mad.wide.u32 (IMAD.WIDE.U32).
template <int VEC>
__device__ __forceinline__ Vec<VEC> gather_row(const float* lane_base, int c, uint32_t pitch_bytes) {
    const float* p;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(c), "r"(pitch_bytes), "l"(lane_base));
    return Vec<VEC>::gather(p);
}

template <bool COHERENT>
__device__ __forceinline__ float4 gather_row4(const float* lane_base, int c, uint32_t pitch_bytes) {
    const float* p;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(c), "r"(pitch_bytes), "l"(lane_base));
    if (!COHERENT) return __ldg(reinterpret_cast<const float4*>(p));
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// One multiply and one add per term instead of a fused multiply-add: TF-CPU's kernel for this op
// (Eigen, stock wheels are built without FMA contraction) rounds the product and the sum
// separately, and so does the CPU checker under tests/ (built with -ffp-contract=off).  With the
// same operation order AND the same roundings, rows that are not split reproduce the checker bit
// for bit (tests/test_gpu_parity.py::test_spmm_unsplit_rows_are_bit_identical_*).
// The kernel is bound by the L1/L2 gather path, not by issue slots: the extra FADD per term costs
// nothing measurable (products shape K=10: 50.5 ms with FMUL+FADD, 51.1 ms with FFMA; profiles/r2/06).
__device__ __forceinline__ float mac(float v, float x, float acc) { return __fadd_rn(__fmul_rn(v, x), acc); }

// Asynchronous global->shared copies (LDGSTS) of one 4-byte word with an L2 evict-first hint;
// src_bytes = 0 writes a zero instead of reading (cp.async zero-fill).
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, int src_bytes, uint64_t pol) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2, %3;" ::"r"(d), "l"(gsrc), "r"(src_bytes),
                 "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// volatile: the two reads of a (col,val) pair (issue phase / math phase) must not be merged
__device__ __forceinline__ int4 lds128(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// A warp owns `rounds * (32/GROUP)` CONSECUTIVE sparse rows and walks them in `rounds` rounds of
// 32/GROUP rows side by side.  The only loads a warp ever WAITS for are the gathers themselves
// (the r1 kernel exposed, per row, the row_ptr load, every (col,val) batch load and the teleport
// row load: ~10 dependent round trips for a 50-entry row where 7 are gathers — profiles/r2/01: the
// bare gather runs the products shape in 3.3 ms, that kernel took 5.07 ms):
//   * row_ptr of all the warp's rows: two coalesced loads up front (lane l holds row l);
//   * the (col,val) batch b+1 — of the same rows, or the first batch of the next round — travels
//     global -> shared as asynchronous copies (cp.async / LDGSTS) into the other half of a
//     per-warp double buffer while batch b is being gathered.  Loads into REGISTERS cannot do this
//     job: a warp's loads return in issue order, so a register prefetch of the (always DRAM-missing)
//     CSR stream issued ahead of a gather batch holds that whole batch back (measured: 59 vs 51 ms);
//   * slots past the end of a row are produced by the copy itself: the column is re-read from the
//     row's last entry and the value is zero-filled (src-size 0), so every batch runs
//     predicate-free for every lane group whatever its degree (0 * x leaves the sum unchanged
//     for finite x, and the padded gathers re-read a feature row that is already in L1);
//   * the teleport row H0[m] is prefetched into L2 at the start of the round.
template <int VEC, int NSLOT, int GROUP, int UNROLL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
spmm_rows_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                 const float* __restrict__ val, const int* __restrict__ row_map,
                 const float* __restrict__ B, int64_t ldb, int n_rows, int long_threshold,
                 int rounds, int piece_ctas, PieceArgs pieces, Epilogue epi) {
    static_assert(UNROLL % 2 == 0 && GROUP % UNROLL == 0, "entries are read back two at a time, whole batches");
    if ((int)blockIdx.x < piece_ctas) {  // CTA-uniform: this CTA works on pieces of split rows
        const int piece = blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
        if (piece < pieces.n_chunks)
            process_piece<VEC, NSLOT, GROUP, (NSLOT >= 4) ? 2 : 4>(row_ptr, col_idx, val, B, ldb, pieces.chunk_row,
                                                                  pieces.chunk_begin, piece, pieces.chunk,
                                                                  pieces.partials, pieces.ldp, epi.F,
                                                                  blockIdx.y * (GROUP * NSLOT * VEC));
        return;
    }
    const int row_cta = blockIdx.x - piece_ctas;
    constexpr int NG = 32 / GROUP;  // rows side by side in a warp
    constexpr int WARPS = THREADS / 32;
    constexpr unsigned FULL = 0xffffffffu;
    // (col,val) pairs of the current and the next batch of this warp; one LDS.128 delivers two
    // entries to every lane of a group (0.5 L1 wavefronts per entry; a shuffle pair costs 2)
    __shared__ __align__(16) int2 cv_smem[WARPS][2][32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g = lane / GROUP;
    const int gl = lane % GROUP;
    const int f_base = blockIdx.y * (GROUP * NSLOT * VEC);
    const int F = epi.F;
    const uint64_t pol = policy_evict_first();
    const uint32_t pitch = (uint32_t)ldb * 4u;

    int fo[NSLOT];
    bool fok[NSLOT];
    const float* Bl[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        fo[s] = f_base + (s * GROUP + gl) * VEC;
        fok[s] = fo[s] < F;  // VEC == 4 implies F % 4 == 0, so the whole slot is in range
        Bl[s] = B + (fok[s] ? fo[s] : 0);
    }

    const int rpw = rounds * NG;  // rows of this warp, <= 32
    const int64_t warp_first = ((int64_t)row_cta * WARPS + warp) * rpw;
    if (warp_first >= n_rows) return;  // warp-uniform; nothing below synchronises across warps
    int rp_lo = 0, rp_hi = 0;          // lane l: row_ptr[warp_first + l], row_ptr[warp_first + l + 1]
    if (lane < rpw && warp_first + lane < n_rows) {
        rp_lo = __ldg(row_ptr + warp_first + lane);
        rp_hi = __ldg(row_ptr + warp_first + lane + 1);
    }
    // rows past the end of the matrix and split rows get deg = 0 and mine = false
    auto round_info = [&](int r, int& start, int& deg, bool& mine, int& maxdeg) {
        const int src = r * NG + g;
        start = __shfl_sync(FULL, rp_lo, src);
        deg = __shfl_sync(FULL, rp_hi, src) - start;
        mine = (warp_first + src < n_rows) && !(long_threshold > 0 && deg > long_threshold);
        if (!mine) deg = 0;
        maxdeg = deg;
        if (NG > 1) {
#pragma unroll
            for (int o = GROUP; o < 32; o <<= 1) maxdeg = max(maxdeg, __shfl_xor_sync(FULL, maxdeg, o));
        }
    };
    // this lane's entry of batch `off` of its group's row -> cv_smem[warp][buf][lane] (asynchronous)
    auto stage = [&](int start, int deg, int off, int buf) {
        int2* dst = &cv_smem[warp][buf][lane];
        if (deg > 0) {
            const int e = min(off + gl, deg - 1);  // clamp: slots past the end repeat the row's last column
            cp_async4(&dst->x, col_idx + start + e, 4, pol);
            cp_async4(&dst->y, val + start + e, (off + gl < deg) ? 4 : 0, pol);
        } else {
            *dst = make_int2(0, 0);  // an idle group: column 0, value 0 (no global address is formed)
        }
        cp_async_commit();
    };

    int start, deg, maxdeg;
    bool mine;
    round_info(0, start, deg, mine, maxdeg);
    int buf = 0;
    stage(start, deg, 0, buf);
    for (int r = 0; r < rounds; ++r) {
        int n_start = 0, n_deg = 0, n_maxdeg = 0;
        bool n_mine = false;
        const bool more = (r + 1 < rounds) && (warp_first + (int64_t)(r + 1) * NG < n_rows);  // warp-uniform
        if (more) round_info(r + 1, n_start, n_deg, n_mine, n_maxdeg);
        const int64_t row = warp_first + r * NG + g;
        const int64_t out_row = (mine && row_map) ? (int64_t)__ldg(row_map + row) : row;
        if (mine && epi.H0 != nullptr) {
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
                if (fok[s]) prefetch_l2(epi.H0 + out_row * epi.ldh + fo[s]);
        }
        Vec<VEC> acc[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[s].v[i] = 0.0f;
        if (maxdeg == 0) {  // nobody reads the batch staged for this round: reuse its buffer for the next round's first
            __syncwarp();
            stage(n_start, n_deg, 0, buf);
        }
        for (int off = 0; off < maxdeg; off += GROUP) {
            __syncwarp();  // every lane is done reading the other buffer (batch before this one)
            if (off + GROUP < maxdeg) stage(start, deg, off + GROUP, buf ^ 1);  // next batch of these rows
            else stage(n_start, n_deg, 0, buf ^ 1);                              // first batch of the next round
            cp_async_wait<1>();  // this lane's copies of the CURRENT batch have landed
            __syncwarp();        // ... and so have everybody else's
            const uint32_t cv_group = smem_u32(&cv_smem[warp][buf][g * GROUP]);
            const int lim = min(GROUP, maxdeg - off);
            for (int j = 0; j < lim; j += UNROLL) {
                // Issue phase: read the columns, issue ALL the batch's gathers back to back.  The values
                // are read again in the math phase instead of being kept: with (col,val) of the whole
                // batch live next to the 4*UNROLL gather registers, ptxas (48-64 registers) interleaved
                // the first FMAs after 3-4 loads, i.e. only 3-4 rows were ever in flight per warp
                // (profiles/r2/05: the first product of a batch held 20 % of all stall samples).
                Vec<VEC> x[UNROLL][NSLOT];
#pragma unroll
                for (int u = 0; u < UNROLL; u += 2) {
                    const int4 e = lds128(cv_group + (j + u) * 8);
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s)
                        if (fok[s]) {
                            x[u][s] = gather_row<VEC>(Bl[s], e.x, pitch);
                            x[u + 1][s] = gather_row<VEC>(Bl[s], e.z, pitch);
                        }
                }
#pragma unroll
                for (int u = 0; u < UNROLL; u += 2) {
                    const int4 e = lds128(cv_group + (j + u) * 8);
                    const float v0 = __int_as_float(e.y), v1 = __int_as_float(e.w);
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s)
                        if (fok[s]) {
#pragma unroll
                            for (int i = 0; i < VEC; ++i) acc[s].v[i] = mac(v0, x[u][s].v[i], acc[s].v[i]);
#pragma unroll
                            for (int i = 0; i < VEC; ++i) acc[s].v[i] = mac(v1, x[u + 1][s].v[i], acc[s].v[i]);
                        }
                }
            }
            buf ^= 1;
        }
        if (deg == 0) {  // an empty (or split) row riding along in a warp: drop whatever the padding produced
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[s].v[i] = 0.0f;
        }
        if (mine) {
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
                if (fok[s]) apply_epilogue<VEC>(epi, out_row, fo[s], acc[s]);
        }
        start = n_start; deg = n_deg; mine = n_mine; maxdeg = n_maxdeg;
    }
    cp_async_wait<0>();  // nothing may still be in flight into this CTA's shared memory at exit
}

// ---------------------------------------------------------------------------------------------
// The float4 fast path (every leading dimension and base pointer 16-byte aligned, F % 4 == 0): the
// same algorithm as spmm_rows_kernel with one float4 slot per lane, written for a small register
// footprint so that ALL `UNROLL` gathers of a batch are in flight before the first product:
//   * no per-slot predicates in the loop: a lane whose slot lies beyond F re-reads slot 0 of the tile
//     (same 32-byte sector as lane 0's read, no extra traffic) and only the final store is predicated;
//   * the next round's (start, degree) are recomputed from the row_ptr registers with two shuffles
//     when they are needed instead of being carried through the loop;
//   * (col,val) pairs are read from shared memory twice (issue phase: columns, math phase: values).
// ---------------------------------------------------------------------------------------------
// COHERENT: the dense operand may have been written earlier in the SAME launch (the persistent
// K-step kernel): its rows are then read with ld.global.cg (L2, always coherent) instead of the
// read-only path.  `cta` = index of this CTA among the row CTAs, `f_tile` = first column of its tile.
template <int GROUP, int UNROLL, int WARPS, bool COHERENT>
__device__ __forceinline__ void rows4_body(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                           const float* __restrict__ val, const int* __restrict__ row_map,
                                           const float* __restrict__ B, int64_t ldb, int n_rows, int long_threshold,
                                           int rounds, int64_t cta, int f_tile, const Epilogue& epi,
                                           int2 (*cv_smem)[2][32]) {
    static_assert(UNROLL % 2 == 0 && GROUP % UNROLL == 0, "entries are read back two at a time, whole batches");
    constexpr int NG = 32 / GROUP;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g = lane / GROUP;
    const int gl = lane % GROUP;
    const int f = f_tile + gl * 4;
    const bool live = f < epi.F;
    const float* Bl = B + (live ? f : f_tile);
    const uint32_t pitch = (uint32_t)ldb * 4u;
    const uint64_t pol = policy_evict_first();

    const int rpw = rounds * NG;  // rows of this warp, <= 32
    const int warp_first = (int)((cta * WARPS + warp) * rpw);  // host: the grid covers < 2^31 rows
    if (warp_first >= n_rows || warp_first < 0) return;
    int rp_lo = 0, rp_hi = 0;  // lane l: row_ptr[warp_first + l], row_ptr[warp_first + l + 1]
    if (lane < rpw && warp_first + lane < n_rows) {
        rp_lo = __ldg(row_ptr + warp_first + lane);
        rp_hi = __ldg(row_ptr + warp_first + lane + 1);
    }
    // (start, degree) of this lane group's row in round r; rows past the end and split rows: degree 0, mine false
    auto round_info = [&](int r, int& start, int& deg, bool& mine) {
        const int src = (r * NG + g) & 31;
        start = __shfl_sync(FULL, rp_lo, src);
        deg = __shfl_sync(FULL, rp_hi, src) - start;
        mine = (r < rounds) && (warp_first + r * NG + g < n_rows) && !(long_threshold > 0 && deg > long_threshold);
        if (!mine) deg = 0;
    };
    auto group_max = [&](int v) {
        if (NG > 1) {
#pragma unroll
            for (int o = GROUP; o < 32; o <<= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
        }
        return v;
    };
    auto stage = [&](int start, int deg, int off, int buf) {
        int2* dst = &cv_smem[warp][buf][lane];
        if (deg > 0) {
            const int e = min(off + gl, deg - 1);  // slots past the end repeat the row's last column with value 0
            cp_async4(&dst->x, col_idx + start + e, 4, pol);
            cp_async4(&dst->y, val + start + e, (off + gl < deg) ? 4 : 0, pol);
        } else {
            *dst = make_int2(0, 0);
        }
        cp_async_commit();
    };

    int buf = 0;
    {
        int s0, d0;
        bool m0;
        round_info(0, s0, d0, m0);
        stage(s0, d0, 0, buf);
    }
    for (int r = 0; r < rounds; ++r) {
        if (warp_first + r * NG >= n_rows) break;  // warp-uniform
        int start, deg;
        bool mine;
        round_info(r, start, deg, mine);
        const int maxdeg = group_max(deg);
        const int row = warp_first + r * NG + g;
        if (mine && live && epi.H0 != nullptr) {
            const int64_t m = row_map ? (int64_t)__ldg(row_map + row) : (int64_t)row;
            prefetch_l2(epi.H0 + m * epi.ldh + f);
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (maxdeg == 0) {  // nobody reads the batch staged for this round: its buffer takes the next round's first batch
            int ns, nd;
            bool nm;
            round_info(r + 1, ns, nd, nm);
            __syncwarp();
            stage(ns, nd, 0, buf);
        }
        for (int off = 0; off < maxdeg; off += GROUP) {
            __syncwarp();  // every lane is done reading the other buffer
            if (off + GROUP < maxdeg) {
                stage(start, deg, off + GROUP, buf ^ 1);
            } else {
                int ns, nd;
                bool nm;
                round_info(r + 1, ns, nd, nm);
                stage(ns, nd, 0, buf ^ 1);
            }
            cp_async_wait<1>();
            __syncwarp();
            const uint32_t cvg = smem_u32(&cv_smem[warp][buf][g * GROUP]);
            const int lim = min(GROUP, maxdeg - off);
            for (int j = 0; j < lim; j += UNROLL) {
                float4 x[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u += 2) {
                    const int4 e = lds128(cvg + (j + u) * 8);
                    x[u] = gather_row4<COHERENT>(Bl, e.x, pitch);
                    x[u + 1] = gather_row4<COHERENT>(Bl, e.z, pitch);
                }
#pragma unroll
                for (int u = 0; u < UNROLL; u += 2) {
                    const int4 e = lds128(cvg + (j + u) * 8);
                    const float v0 = __int_as_float(e.y), v1 = __int_as_float(e.w);
                    acc.x = mac(v0, x[u].x, acc.x); acc.y = mac(v0, x[u].y, acc.y);
                    acc.z = mac(v0, x[u].z, acc.z); acc.w = mac(v0, x[u].w, acc.w);
                    acc.x = mac(v1, x[u + 1].x, acc.x); acc.y = mac(v1, x[u + 1].y, acc.y);
                    acc.z = mac(v1, x[u + 1].z, acc.z); acc.w = mac(v1, x[u + 1].w, acc.w);
                }
            }
            buf ^= 1;
        }
        if (mine && live) {
            if (deg == 0) acc = make_float4(0.f, 0.f, 0.f, 0.f);  // padding of a longer neighbour row in this warp
            const int64_t m = row_map ? (int64_t)__ldg(row_map + row) : (int64_t)row;
            apply_epilogue<4>(epi, m, f, Vec<4>{{acc.x, acc.y, acc.z, acc.w}});
        }
    }
    cp_async_wait<0>();  // nothing may still be in flight into this CTA's shared memory at exit
}


// Grid layout along x: [ push CTAs | piece CTAs | row CTAs ].  The push CTAs (sharded runs only) send the
// rows of the dense operand B that the peers need straight into their halo buffers over NVLink and
// signal completion (push.cuh) while the rest of the grid computes: being the lowest block indices
// they are dispatched first, so the exchange overlaps the whole launch without a second stream.
template <int GROUP, int UNROLL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
spmm_rows4_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                  const float* __restrict__ val, const int* __restrict__ row_map,
                  const float* __restrict__ B, int64_t ldb, int n_rows, int long_threshold,
                  int rounds, int piece_ctas, PieceArgs pieces, Epilogue epi, PushArgs push) {
    if ((int)blockIdx.x < push.n_ctas) {  // CTA-uniform
        if (blockIdx.y == 0) {
            halo_push_body<4, GROUP, 4>(B, ldb, push.send_idx, push.send_off, push.peer_base, push.peer_row0, push.n_peers,
                                        push.n_send, push.rotate, push.ldo, epi.F, (int)blockIdx.x, push.n_ctas);
            push_signal_tail(push.sg, push.n_peers, push.n_ctas);
        }
        return;
    }
    const int bx = (int)blockIdx.x - push.n_ctas;
    if (bx < piece_ctas) {  // CTA-uniform: this CTA works on pieces of split rows
        const int piece = bx * (THREADS / 32) + (threadIdx.x >> 5);
        if (piece < pieces.n_chunks)
            process_piece<4, 1, GROUP, 4>(row_ptr, col_idx, val, B, ldb, pieces.chunk_row, pieces.chunk_begin, piece,
                                          pieces.chunk, pieces.partials, pieces.ldp, epi.F, blockIdx.y * (GROUP * 4));
        return;
    }
    __shared__ __align__(16) int2 cv_smem[THREADS / 32][2][32];
    rows4_body<GROUP, UNROLL, THREADS / 32, false>(row_ptr, col_idx, val, row_map, B, ldb, n_rows, long_threshold, rounds,
                                                  (int64_t)bx - piece_ctas, blockIdx.y * (GROUP * 4), epi, cv_smem);
}

// ---------------------------------------------------------------------------------------------
// Persistent K-step kernel for graphs whose whole step is ONE wave of CTAs (Cora, PubMed ...): the K
// PPRIteration layers (filter.py:34-35 under layered.py:52-55) run inside one cooperative launch
// with a grid-wide barrier between steps instead of K launches — those shapes are bound by launch
// latency (r1: Cora 6.7 us per step for a 0.4 MB working set), not by memory.  Step k reads
// buffer src_k and writes dst_k exactly as gnntf_appnp_propagate_f32 does (dst_{K-1} = H_out).
// Buffers written inside the launch are gathered with ld.global.cg; no row is split (the host
// takes this path only when the long-row plan is empty).
// ---------------------------------------------------------------------------------------------
template <int GROUP, int UNROLL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
appnp_persistent_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                        const float* __restrict__ val, const float* __restrict__ H0, float* __restrict__ H_out,
                        float* __restrict__ scratch, int64_t ld, int n_rows, int rounds, int K, Epilogue epi) {
    __shared__ __align__(16) int2 cv_smem[THREADS / 32][2][32];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const float* src = H0;
    for (int k = 0; k < K; ++k) {
        float* dst = ((K - 1 - k) % 2 == 0) ? H_out : scratch;
        Epilogue e = epi;
        e.C = dst;
        if (k == 0)
            rows4_body<GROUP, UNROLL, THREADS / 32, false>(row_ptr, col_idx, val, nullptr, src, ld, n_rows, 0, rounds,
                                                          (int64_t)blockIdx.x, blockIdx.y * (GROUP * 4), e, cv_smem);
        else
            rows4_body<GROUP, UNROLL, THREADS / 32, true>(row_ptr, col_idx, val, nullptr, src, ld, n_rows, 0, rounds,
                                                         (int64_t)blockIdx.x, blockIdx.y * (GROUP * 4), e, cv_smem);
        src = dst;
        if (k + 1 < K) grid.sync();
    }
}

// Long rows, phase 2: one CTA per long row; thread t owns feature t (strided), sums the row's
// pieces in piece order, then applies the epilogue.
__global__ void __launch_bounds__(128)
spmm_long_reduce_kernel(const int* __restrict__ long_row, const int* __restrict__ long_first_chunk,
                        const int* __restrict__ long_n_chunks, const int* __restrict__ row_map,
                        const float* __restrict__ partials, int ldp, Epilogue epi) {
    const int li = blockIdx.x;
    const int csr_row = __ldg(long_row + li);
    const int64_t row = row_map ? (int64_t)__ldg(row_map + csr_row) : (int64_t)csr_row;
    const int first = __ldg(long_first_chunk + li);
    const int cnt = __ldg(long_n_chunks + li);
    for (int f = threadIdx.x; f < epi.F; f += blockDim.x) {
        const float* p = partials + (int64_t)first * ldp + f;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // 4 independent chains, fixed assignment
        int k = 0;
        for (; k + 4 <= cnt; k += 4) {
            a0 += ld_stream(p + (int64_t)(k + 0) * ldp);
            a1 += ld_stream(p + (int64_t)(k + 1) * ldp);
            a2 += ld_stream(p + (int64_t)(k + 2) * ldp);
            a3 += ld_stream(p + (int64_t)(k + 3) * ldp);
        }
        for (; k < cnt; ++k) a0 += ld_stream(p + (int64_t)k * ldp);
        Vec<1> acc{{(a0 + a1) + (a2 + a3)}};
        apply_epilogue<1>(epi, row, f, acc);
    }
}

// ---------------------------------------------------------------------------------------------
// Host-side dispatch
// ---------------------------------------------------------------------------------------------
// Rounds per warp.  Deep grids: 8 (a warp's row_ptr values live in one register per lane, so
// rounds * rows-side-by-side <= 32).  Grids of only a few waves (arxiv: ~4) lose up to a whole wave
// to quantisation, so among the candidates the one whose last wave is fullest wins; small graphs
// (Cora: 43 row blocks) keep one round so that every SM gets work.
static int pick_rounds(int64_t n_rows, int rows_per_round, int ng, int ctas_per_sm) {
    const int64_t row_blocks = ceil_div(n_rows, rows_per_round);
    const int64_t slots = (int64_t)kNumSMs * ctas_per_sm;
    const int cap = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(8, 32 / ng), row_blocks / (slots * 4)));
    int best = cap;
    double best_fill = 0.0;
    for (int r = cap; r >= std::max(1, cap / 2); --r) {
        const double waves = (double)ceil_div(row_blocks, r) / (double)slots;
        const double fill = waves / std::ceil(waves);
        if (fill > best_fill + 0.02) { best_fill = fill; best = r; }
    }
    return best;
}

template <int VEC, int NSLOT, int GROUP, int THREADS, int MINB, int UNROLL_>
static int launch_cfg_t(const gnntf_csr_t* A, const float* B, int64_t ldb, const Epilogue& epi,
                        cudaStream_t st) {
    constexpr int UNROLL = UNROLL_ > 0 ? UNROLL_ : ((NSLOT >= 4) ? 2 : (NSLOT == 2 ? 4 : (GROUP >= 8 ? 8 : 4)));
    constexpr int NG = 32 / GROUP;
    constexpr int ROWS_PER_ROUND = (THREADS / 32) * NG;  // rows of one CTA per round
    const int F = epi.F;
    const int tile = GROUP * NSLOT * VEC;
    const unsigned gy = (unsigned)ceil_div(F, tile);
    const int thr = (A->n_long > 0) ? A->long_threshold : 0;
    if (A->n_rows > 0) {
        // several rounds per warp once the grid is deep enough to keep every SM busy for many
        // waves; small graphs (Cora: 43 row blocks) keep one round per warp.  A warp holds the
        // row_ptr of all its rows in one register per lane: rounds * NG <= 32.
        const int rounds = pick_rounds(A->n_rows, ROWS_PER_ROUND, NG, MINB);
        // the pieces of split rows ride in the same grid, ahead of the ordinary rows
        PieceArgs pieces{};
        int piece_ctas = 0;
        const int ldp = (int)round_up(F, 4);
        if (A->n_long > 0) {
            pieces = PieceArgs{A->chunk_row, A->chunk_begin, A->n_chunks, A->chunk, A->partials, ldp};
            piece_ctas = (int)ceil_div(A->n_chunks, THREADS / 32);
        }
        dim3 grid((unsigned)(piece_ctas + ceil_div(A->n_rows, (int64_t)ROWS_PER_ROUND * rounds)), gy);
        spmm_rows_kernel<VEC, NSLOT, GROUP, UNROLL, THREADS, MINB><<<grid, THREADS, 0, st>>>(
            A->row_ptr, A->col_idx, A->val, A->row_map, B, ldb, (int)A->n_rows, thr, rounds, piece_ctas, pieces, epi);
        GNNTF_LAUNCH_CHECK();
        if (A->n_long > 0) {
            spmm_long_reduce_kernel<<<A->n_long, 128, 0, st>>>(A->long_row, A->long_first_chunk,
                                                              A->long_n_chunks, A->row_map, A->partials,
                                                              ldp, epi);
            GNNTF_LAUNCH_CHECK();
        }
    }
    return GNNTF_OK;
}

template <int VEC, int NSLOT, int GROUP>
static int launch_cfg(const gnntf_csr_t* A, const float* B, int64_t ldb, const Epilogue& epi,
                      cudaStream_t st) {
    return launch_cfg_t<VEC, NSLOT, GROUP, 256, 5, 0>(A, B, ldb, epi, st);
}

// float4 fast path: spmm_rows4_kernel
template <int GROUP, int THREADS, int MINB>
static int launch_rows4_t(const gnntf_csr_t* A, const float* B, int64_t ldb, const Epilogue& epi, cudaStream_t st,
                          const PushArgs& push) {
    constexpr int UNROLL = (GROUP >= 8) ? 8 : 4;
    constexpr int NG = 32 / GROUP;
    constexpr int ROWS_PER_ROUND = (THREADS / 32) * NG;
    const int F = epi.F;
    const unsigned gy = (unsigned)ceil_div(F, GROUP * 4);
    const int thr = (A->n_long > 0) ? A->long_threshold : 0;
    if (A->n_rows == 0 && push.n_ctas == 0) return GNNTF_OK;
    const int rounds = pick_rounds(std::max<int64_t>(A->n_rows, 1), ROWS_PER_ROUND, NG, MINB);
    PieceArgs pieces{};
    int piece_ctas = 0;
    const int ldp = (int)round_up(F, 4);
    if (A->n_long > 0) {
        pieces = PieceArgs{A->chunk_row, A->chunk_begin, A->n_chunks, A->chunk, A->partials, ldp};
        piece_ctas = (int)ceil_div(A->n_chunks, THREADS / 32);
    }
    dim3 grid((unsigned)(push.n_ctas + piece_ctas + ceil_div(A->n_rows, (int64_t)ROWS_PER_ROUND * rounds)), gy);
    spmm_rows4_kernel<GROUP, UNROLL, THREADS, MINB><<<grid, THREADS, 0, st>>>(
        A->row_ptr, A->col_idx, A->val, A->row_map, B, ldb, (int)A->n_rows, thr, rounds, piece_ctas, pieces, epi, push);
    GNNTF_LAUNCH_CHECK();
    if (A->n_long > 0) {
        spmm_long_reduce_kernel<<<A->n_long, 128, 0, st>>>(A->long_row, A->long_first_chunk, A->long_n_chunks,
                                                          A->row_map, A->partials, ldp, epi);
        GNNTF_LAUNCH_CHECK();
    }
    return GNNTF_OK;
}

template <int GROUP>
static int launch_rows4(const gnntf_csr_t* A, const float* B, int64_t ldb, const Epilogue& epi, cudaStream_t st,
                        const PushArgs& push) {
    // 256 threads x 5 CTAs per SM (48 registers): A/B against 256x4, 256x3, 192x6, 128x8 — all within 2 %
    // on the arxiv and products shapes (profiles/r2/06)
    return launch_rows4_t<GROUP, 256, 5>(A, B, ldb, epi, st, push);
}



static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Persistent K-step launch (appnp_persistent_kernel).  Returns GNNTF_OK and sets *taken when the
// shape qualifies and the launch was enqueued; leaves *taken false otherwise (the caller then
// issues K ordinary launches).
template <int GROUP>
static int try_persistent(const gnntf_csr_t* A, const float* H0, float* H_out, float* scratch, int64_t ld,
                          const Epilogue& epi, int K, cudaStream_t st, bool* taken, bool single_round_only = false) {
    constexpr int THREADS = 256, MINB = 5;
    constexpr int UNROLL = (GROUP >= 8) ? 8 : 4;
    constexpr int NG = 32 / GROUP;
    auto kern = appnp_persistent_kernel<GROUP, UNROLL, THREADS, MINB>;
    static int slots = -1;  // co-resident CTAs of this instantiation on the current device (0 = no cooperative launch)
    if (slots < 0) {
        int dev = 0, coop = 0, per_sm = 0, sms = 0;
        GNNTF_CUDA_TRY(cudaGetDevice(&dev));
        GNNTF_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        GNNTF_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        GNNTF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, 0));
        slots = coop ? per_sm * sms : 0;
    }
    const int64_t gy = ceil_div(epi.F, GROUP * 4);
    int rounds = 1;
    while (rounds < 32 / NG && ceil_div(A->n_rows, (int64_t)(THREADS / 32) * NG * rounds) * gy > slots) rounds *= 2;
    const int64_t gx = ceil_div(A->n_rows, (int64_t)(THREADS / 32) * NG * rounds);
    if (slots == 0 || gx * gy > slots) return GNNTF_OK;
    if (single_round_only && rounds > 1) return GNNTF_OK;
    const int* row_ptr = A->row_ptr;
    const int* col_idx = A->col_idx;
    const float* val = A->val;
    int n_rows = (int)A->n_rows;
    Epilogue e = epi;
    void* args[] = {&row_ptr, &col_idx, &val, &H0, &H_out, &scratch, &ld, &n_rows, &rounds, &K, &e};
    GNNTF_CUDA_TRY(cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)gx, (unsigned)gy), dim3(THREADS), args, 0, st));
    *taken = true;
    return GNNTF_OK;
}

int spmm_persistent_propagate(const gnntf_csr_t* A, const float* H0, float* H_out, float* scratch, int64_t ld,
                              int64_t F, double alpha, int K, cudaStream_t st, bool* taken) {
    *taken = false;
    if (K < 2 || A->n_long > 0 || A->row_map != nullptr || A->n_rows <= 0 || F <= 0 || scratch == nullptr) return GNNTF_OK;
    if (F % 4 != 0 || ld % 4 != 0 || ld >= (1LL << 30)) return GNNTF_OK;
    if (!aligned16(H0) || !aligned16(H_out) || !aligned16(scratch)) return GNNTF_OK;
    Epilogue e{};
    e.s = (float)(1.0 - alpha);
    e.H0 = H0;
    e.ldh = ld;
    e.t = (float)alpha;
    e.act = GNNTF_ACT_IDENTITY;
    e.ldc = ld;
    e.F = (int)F;
    const int64_t slots = F / 4;
    // These shapes are bound by latency, not by lanes: a lane group WIDER than the row needs (idle lanes) puts fewer
    // rows on a warp and spreads the graph over more SMs — Cora, F = 8: 43 CTAs with 4-lane groups, 339 with a warp
    // per row (measured: the F = 64 mapping ran the same graph in 42 us where the F = 8 mapping took 55 us).
    // Taken while every warp still needs a single round and the grid stays co-resident.
    int rc = GNNTF_OK;
    if (slots <= 16) {
        rc = try_persistent<32>(A, H0, H_out, scratch, ld, e, K, st, taken, true);
        if (rc != GNNTF_OK || *taken) return rc;
    }
    if (slots <= 8) {
        rc = try_persistent<16>(A, H0, H_out, scratch, ld, e, K, st, taken, true);
        if (rc != GNNTF_OK || *taken) return rc;
    }
    if (slots <= 4) {
        rc = try_persistent<8>(A, H0, H_out, scratch, ld, e, K, st, taken, true);
        if (rc != GNNTF_OK || *taken) return rc;
        return try_persistent<4>(A, H0, H_out, scratch, ld, e, K, st, taken);
    }
    if (slots <= 8) return try_persistent<8>(A, H0, H_out, scratch, ld, e, K, st, taken);
    if (slots <= 16) return try_persistent<16>(A, H0, H_out, scratch, ld, e, K, st, taken);
    return try_persistent<32>(A, H0, H_out, scratch, ld, e, K, st, taken);
}

int validate_csr(const gnntf_csr_t* A) {
    if (A == nullptr) return GNNTF_E_NULL;
    if (A->n_rows < 0 || A->nnz < 0 || A->nnz > 0x7fffffffLL || A->n_rows > 0x7ffffffeLL) return GNNTF_E_SIZE;
    if (A->n_rows > 0 && A->row_ptr == nullptr) return GNNTF_E_NULL;
    if (A->nnz > 0 && (A->col_idx == nullptr || A->val == nullptr)) return GNNTF_E_NULL;
    if (A->n_long < 0 || A->n_chunks < 0) return GNNTF_E_SIZE;
    if (A->n_long > 0) {
        if (A->long_row == nullptr || A->long_first_chunk == nullptr || A->long_n_chunks == nullptr ||
            A->chunk_row == nullptr || A->chunk_begin == nullptr || A->partials == nullptr)
            return GNNTF_E_NULL;
        if (A->long_threshold <= 0 || A->chunk <= 0) return GNNTF_E_SIZE;
    }
    return GNNTF_OK;
}

// C (and ACC) = epilogue(A·B).  Chooses the lane mapping from F and the alignment of every
// operand; see the file header.
int spmm_dispatch(const gnntf_csr_t* A, const float* B, int64_t ldb, Epilogue epi, cudaStream_t st,
                  const PushArgs* push, bool* push_done) {
    if (push_done) *push_done = false;
    int rc = validate_csr(A);
    if (rc != GNNTF_OK) return rc;
    const int64_t F = epi.F;
    if (F < 0 || F > 0x7fffffff) return GNNTF_E_SIZE;
    if (A->n_rows == 0 || F == 0) return GNNTF_OK;
    if (B == nullptr) return GNNTF_E_NULL;
    if (ldb >= (1LL << 30)) return GNNTF_E_SIZE;  // row pitch in bytes must fit 32 bits
    if (ldb < F || (epi.C && epi.ldc < F) || (epi.H0 && epi.ldh < F) || (epi.ACC && epi.ldacc < F))
        return GNNTF_E_SIZE;
    epi.B = B;
    epi.ldb = ldb;

    bool v4 = (F % 4 == 0) && (ldb % 4 == 0) && aligned16(B);
    if (epi.C) v4 = v4 && (epi.ldc % 4 == 0) && aligned16(epi.C);
    if (epi.H0) v4 = v4 && (epi.ldh % 4 == 0) && aligned16(epi.H0);
    if (epi.ACC) v4 = v4 && (epi.ldacc % 4 == 0) && aligned16(epi.ACC);
    if (epi.keep) v4 = false;  // byte mask rows are F-strided: keep the scalar path

    if (v4) {
        PushArgs pa{};
        if (push != nullptr && push->n_ctas > 0 && push->ldo % 4 == 0) {  // the push rides in this launch
            pa = *push;
            if (push_done) *push_done = true;
        }
        const int64_t slots = F / 4;
        if (slots <= 4) return launch_rows4<4>(A, B, ldb, epi, st, pa);
        if (slots <= 8) return launch_rows4<8>(A, B, ldb, epi, st, pa);
        if (slots <= 16) return launch_rows4<16>(A, B, ldb, epi, st, pa);
        // wider rows: 128-float tiles over grid.y (multi-slot mappings spilled and lost, profiles/r1/09)
        return launch_rows4<32>(A, B, ldb, epi, st, pa);
    }
    if (F <= 4) return launch_cfg<1, 1, 4>(A, B, ldb, epi, st);
    if (F <= 8) return launch_cfg<1, 1, 8>(A, B, ldb, epi, st);
    if (F <= 16) return launch_cfg<1, 1, 16>(A, B, ldb, epi, st);
    if (F <= 32) return launch_cfg<1, 1, 32>(A, B, ldb, epi, st);
    if (F <= 64) return launch_cfg<1, 2, 32>(A, B, ldb, epi, st);
    return launch_cfg<1, 4, 32>(A, B, ldb, epi, st);  // tiles of 128 floats over grid.y
}

// ---------------------------------------------------------------------------------------------
// Long-row plan
// ---------------------------------------------------------------------------------------------
__global__ void plan_count_kernel(const int* __restrict__ row_ptr, int n_rows, int thr, int chunk,
                                  int* __restrict__ counts) {
    int nl = 0, nc = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int deg = row_ptr[r + 1] - row_ptr[r];
        if (deg > thr) {
            nl += 1;
            nc += (deg + chunk - 1) / chunk;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nl += __shfl_xor_sync(0xffffffffu, nl, o);
        nc += __shfl_xor_sync(0xffffffffu, nc, o);
    }
    if ((threadIdx.x & 31) == 0 && nl > 0) {
        atomicAdd(counts + 0, nl);
        atomicAdd(counts + 1, nc);
    }
}

// Long rows are rare; one thread claims a contiguous range of pieces for its row.  The order in
// which rows claim ranges is arbitrary but never changes a result (pieces of one row stay
// contiguous and ordered).
__global__ void plan_fill_kernel(const int* __restrict__ row_ptr, int n_rows, int thr, int chunk,
                                 int* __restrict__ counters, int* __restrict__ long_row,
                                 int* __restrict__ long_first_chunk, int* __restrict__ long_n_chunks,
                                 int* __restrict__ chunk_row, int* __restrict__ chunk_begin) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int start = row_ptr[r];
        const int deg = row_ptr[r + 1] - start;
        if (deg > thr) {
            const int nch = (deg + chunk - 1) / chunk;
            const int li = atomicAdd(counters + 0, 1);
            const int first = atomicAdd(counters + 1, nch);
            long_row[li] = (int)r;
            long_first_chunk[li] = first;
            long_n_chunks[li] = nch;
            for (int j = 0; j < nch; ++j) {
                chunk_row[first + j] = (int)r;
                chunk_begin[first + j] = start + j * chunk;
            }
        }
    }
}

}  // namespace gnntf

using namespace gnntf;

extern "C" int gnntf_spmm_plan_count(const int32_t* row_ptr, int64_t n_rows, int32_t long_threshold,
                                     int32_t chunk, int32_t* counts, void* stream) {
    if (counts == nullptr || (n_rows > 0 && row_ptr == nullptr)) return GNNTF_E_NULL;
    if (n_rows < 0 || n_rows > 0x7ffffffeLL || long_threshold <= 0 || chunk <= 0) return GNNTF_E_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    GNNTF_CUDA_TRY(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st));
    if (n_rows == 0) return GNNTF_OK;
    const int grid = (int)std::min<int64_t>(ceil_div(n_rows, 256), (int64_t)kNumSMs * 8);
    plan_count_kernel<<<grid, 256, 0, st>>>(row_ptr, (int)n_rows, long_threshold, chunk, counts);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_spmm_plan_fill(const int32_t* row_ptr, int64_t n_rows, int32_t long_threshold,
                                    int32_t chunk, int32_t* counters_ws, int32_t* long_row,
                                    int32_t* long_first_chunk, int32_t* long_n_chunks,
                                    int32_t* chunk_row, int32_t* chunk_begin, void* stream) {
    if (counters_ws == nullptr || (n_rows > 0 && row_ptr == nullptr)) return GNNTF_E_NULL;
    if (n_rows < 0 || n_rows > 0x7ffffffeLL || long_threshold <= 0 || chunk <= 0) return GNNTF_E_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    GNNTF_CUDA_TRY(cudaMemsetAsync(counters_ws, 0, 2 * sizeof(int32_t), st));
    if (n_rows == 0) return GNNTF_OK;
    const int grid = (int)std::min<int64_t>(ceil_div(n_rows, 256), (int64_t)kNumSMs * 8);
    plan_fill_kernel<<<grid, 256, 0, st>>>(row_ptr, (int)n_rows, long_threshold, chunk, counters_ws,
                                           long_row, long_first_chunk, long_n_chunks, chunk_row,
                                           chunk_begin);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_spmm_f32(const gnntf_csr_t* A, const float* B, int64_t ldb, float* C,
                              int64_t ldc, int64_t F, void* stream) {
    if (C == nullptr && A != nullptr && A->n_rows > 0 && F > 0) return GNNTF_E_NULL;
    Epilogue e{};
    e.s = 1.0f;
    e.C = C;
    e.ldc = ldc;
    e.F = (int)F;
    e.act = GNNTF_ACT_IDENTITY;
    if (F < 0 || F > 0x7fffffff) return GNNTF_E_SIZE;
    return spmm_dispatch(A, B, ldb, e, (cudaStream_t)stream);
}

extern "C" int gnntf_spmm_acc_f32(const gnntf_csr_t* A, const float* B, int64_t ldb, float* C,
                                  int64_t ldc, int64_t F, double scale, void* stream) {
    if (C == nullptr && A != nullptr && A->n_rows > 0 && F > 0) return GNNTF_E_NULL;
    if (F < 0 || F > 0x7fffffff) return GNNTF_E_SIZE;
    Epilogue e{};
    e.s = (float)scale;
    e.act = GNNTF_ACT_IDENTITY;
    e.F = (int)F;
    e.C = nullptr;   // the product only feeds the accumulator
    e.ACC = C;
    e.ldacc = ldc;
    e.u = 0.0f;
    e.w = 1.0f;
    e.acc_init = 0;
    return spmm_dispatch(A, B, ldb, e, (cudaStream_t)stream);
}
