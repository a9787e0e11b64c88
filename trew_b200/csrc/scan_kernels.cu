// Hand-written CUDA (sm_100a) for TREW's per-read tandem-repeat scan-and-count.
// See scan_kernels.cuh for the two-kernel structure and DESIGN.md for the derivation.
//
// Reference semantics restated here (paths into /root/reference):
//   codes / canonical rotation / reverse complement   src/kmer.cpp:7-78, 1815-1867
//   k_mer_check(_128)  (SCAN)                          src/kmer.cpp:2144-2547
//   k_mer_target(_128) (TARGET)                        src/kmer.cpp:1894-2142
//   buffer_task / _pair / _long (routing)              src/kmer.cpp:80-985
//
// Key facts used (proved in DESIGN.md):
//   L1  two adjacent valid k-windows i, i+1 are in the same rotation class iff base[i] == base[i+k];
//       so classes are unions of maximal match-bit runs and one canonicalisation per run suffices.
//   L2  any rotation-invariant signature partitions windows more coarsely than rotation classes, so the
//       largest signature bucket bounds the largest class from above.  The signature used is the parity
//       of the hi-bit count, lo-bit count (and, second level, A count) of the window, computed for all
//       window positions at once from prefix-XOR bit-planes: sig_k = (P >> k) ^ P.
#include "scan_kernels.cuh"

#include <cstdio>

namespace trew {

typedef unsigned int u32;
typedef unsigned long long u64;
typedef unsigned __int128 u128;

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ u32 low_mask(int bits) {  // bits in [0, 32]
    return bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
}

__device__ __forceinline__ u32 prefix_xor32(u32 x) {  // bit j = xor of bits 0..j
    x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;
    return x;
}

__device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// ------------------------------------------------------------------------------------------------
// device count table: (table, k, seq128) -> count, open addressing, linear probing
// (stands in for the per-worker ResultMap objects, src/kmer.h:79-81)
// ------------------------------------------------------------------------------------------------

__device__ __noinline__ void table_add(const DevCfg& cfg, u32 meta, u64 lo, u64 hi, u64 cnt) {
    u64 h = mix64(lo ^ mix64(hi + 0x9e3779b97f4a7c15ULL * (u64)(meta + 1)));
    u32 i = (u32)h & cfg.slot_mask;
    for (u32 probe = 0; probe <= cfg.slot_mask; probe++, i = (i + 1) & cfg.slot_mask) {
        Slot* s = cfg.slots + i;
        u32 st = atomicCAS(&s->state, 0u, 1u);
        if (st == 0u) {
            s->seq_lo = lo; s->seq_hi = hi; s->meta = meta;
            __threadfence();
            atomicExch(&s->state, 2u);
            atomicAdd(&s->count, cnt);
            return;
        }
        while (st == 1u) st = *(volatile u32*)&s->state;
        __threadfence();
        if (__ldcg(&s->meta) == meta && __ldcg(&s->seq_lo) == lo && __ldcg(&s->seq_hi) == hi) {
            atomicAdd(&s->count, cnt);
            return;
        }
    }
    atomicExch(cfg.error_flag, 3u);  // TREW_ERR_TABLE_FULL
}

__global__ void compact_kernel(const Slot* __restrict__ slots, u32 n_slots, u32* __restrict__ d_meta,
                               u64* __restrict__ d_seq, u64* __restrict__ d_count, u32* __restrict__ d_n) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) {
        Slot s = slots[i];
        bool used = s.state == 2u && s.count != 0;
        u32 m = __ballot_sync(0xffffffffu, used);
        if (m) {
            u32 base = 0;
            if (lane_id() == (u32)(__ffs(m) - 1)) base = atomicAdd(d_n, (u32)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (used) {
                u32 o = base + __popc(m & ((1u << lane_id()) - 1u));
                d_meta[o] = s.meta; d_seq[2 * (size_t)o] = s.seq_lo; d_seq[2 * (size_t)o + 1] = s.seq_hi; d_count[o] = s.count;
            }
        }
    }
}

void launch_compact(const Slot* slots, unsigned int n_slots, unsigned int* d_meta, unsigned long long* d_seq,
                    unsigned long long* d_count, unsigned int* d_n, cudaStream_t stream) {
    // n_slots is a power of two >= 1024, so every warp iterates the same number of times
    compact_kernel<<<592, 256, 0, stream>>>(slots, n_slots, d_meta, d_seq, d_count, d_n);
}

// ------------------------------------------------------------------------------------------------
// synthetic batch generator (benchmark tooling; mirrored by trew_b200/synth.py:device_mirror)
// ------------------------------------------------------------------------------------------------

__host__ __device__ inline u32 synth_hash(u64 seed, u64 a, u64 b) {
    u64 x = seed + 0x9e3779b97f4a7c15ULL * (a + 1) + 0xbf58476d1ce4e5b9ULL * (b + 1);
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31;
    return (u32)(x >> 32);
}

__global__ void synth_kernel(u64 seed, u32 n_reads, u32 L, u32 tel_thr, u32 half_thr, u32 n_thr, u32 sub_thr,
                             u32* __restrict__ bit_off, u32* __restrict__ hi, u32* __restrict__ lo, u32* __restrict__ val,
                             size_t plane_words) {
    const u64 total = (u64)n_reads * L;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_reads; i += stride) bit_off[i] = (u32)(i * L);
    // codes T=0 G=1 C=2 A=3; TTAGGG and its reverse complement CCCTAA as code strings
    const int unit_f[6] = {0, 0, 3, 1, 1, 1};
    const int unit_r[6] = {2, 2, 2, 0, 3, 3};
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < plane_words; w += stride) {
        u32 h = 0, l = 0, v = 0;
        for (int b = 0; b < 32; b++) {
            u64 pos = (u64)w * 32 + b;
            if (pos >= total) break;
            u64 r = pos / L; u32 j = (u32)(pos % L);
            u32 kind = synth_hash(seed, r, 0xffffffffULL);
            u32 aux = synth_hash(seed, r, 0xfffffffeULL);
            u32 code = synth_hash(seed, r, j) & 3u;
            bool tel = kind < tel_thr;
            bool halfk = !tel && kind < tel_thr + half_thr;
            if (halfk) tel = ((aux >> 8) & 1u) ? (j < L / 2) : (j >= L / 2);
            if (tel) {
                u32 phase = aux % 6u;
                bool rc = (aux >> 4) & 1u;
                u32 idx = (j + phase) % 6u;
                u32 c = rc ? unit_r[idx] : unit_f[idx];
                u32 sh = synth_hash(seed ^ 0x5555555555555555ULL, r, j);
                code = sh < sub_thr ? (sh >> 3) & 3u : c;   // note: sh < sub_thr keeps (sh >> 3) & 3 uniform enough
            }
            bool inval = synth_hash(seed ^ 0xaaaaaaaaaaaaaaaaULL, r, j) < n_thr;
            if (!inval) { v |= 1u << b; h |= (code >> 1) << b; l |= (code & 1u) << b; }
        }
        hi[w] = h; lo[w] = l; val[w] = v;
    }
}

void launch_synth(unsigned long long seed, unsigned int n_reads, unsigned int read_len, unsigned int tel_thr,
                  unsigned int half_thr, unsigned int n_thr, unsigned int sub_thr, unsigned int* bit_off, unsigned int* hi,
                  unsigned int* lo, unsigned int* val, size_t plane_words, cudaStream_t stream) {
    synth_kernel<<<1184, 256, 0, stream>>>(seed, n_reads, read_len, tel_thr, half_thr, n_thr, sub_thr, bit_off, hi, lo, val,
                                            plane_words);
}

// ------------------------------------------------------------------------------------------------
// probe windows: the first windows the reference's routing scans for a unit.  If none of them can
// produce a target k, the unit emits nothing ("Why the filter is sound").
// ------------------------------------------------------------------------------------------------

struct Probe { u32 pos; int wl, k0, k1; };

__device__ __forceinline__ int unit_probes(const DevCfg& c, const DevBatch& b, u32 u, Probe* p) {
    const int MINM = c.min_mer, MAXM = c.max_mer;
    int np = 0;
    if (c.mode == 0) {  // buffer_task, src/kmer.cpp:111-171
        u32 b0 = __ldg(b.bit_off + u);
        int n = (int)(__ldg(b.bit_off + u + 1) - b0);
        if (n < 2 * MINM) return 0;
        if (n >= 4 * MINM) {
            int kmax = min(n / 4, MAXM);
            p[np++] = Probe{b0, n / 2, MINM, kmax};
            p[np++] = Probe{b0 + (u32)(n - (n + 1) / 2), (n + 1) / 2, MINM, kmax};
        }
        if (4 * MAXM > n) p[np++] = Probe{b0, n, max(n / 4 + 1, MINM), min(n / 2, MAXM)};
    } else if (c.mode == 1) {  // buffer_task_pair, src/kmer.cpp:322-505
        u32 a0 = __ldg(b.bit_off + 2 * u), a1 = __ldg(b.bit_off + 2 * u + 1), a2 = __ldg(b.bit_off + 2 * u + 2);
        int n1 = (int)(a1 - a0), n2 = (int)(a2 - a1);
        int n = min(n1, n2);
        if (n < 2 * MINM) return 0;
        if (n >= 4 * MINM) {
            int kmax = min(n / 4, MAXM);
            p[np++] = Probe{a0, n1 / 2, MINM, kmax};  // segment 1: forward walk starts here
            p[np++] = Probe{a1, n2 / 2, MINM, kmax};  // segment 4: backward walk starts here
        }
        if (4 * MAXM > n) {
            int lo = max(n / 4 + 1, MINM), hi = min(n / 2, MAXM);
            p[np++] = Probe{a0, n1, lo, hi};
            p[np++] = Probe{a1, n2, lo, hi};
        }
    } else {  // buffer_task_long, src/kmer.cpp:785-856
        u32 b0 = __ldg(b.bit_off + u);
        int n = (int)(__ldg(b.bit_off + u + 1) - b0);
        const int SL = c.slice_len;
        if (n < SL) return 0;
        int snum = n / SL, mid = (snum + 1) / 2, bonus = n % SL;
        int l1 = SL + (1 == mid ? bonus : 0);
        p[np++] = Probe{b0, l1, MINM, MAXM};
        if (snum > 1) {
            int l2 = SL + (snum == mid ? bonus : 0);
            p[np++] = Probe{b0 + (u32)(n - l2), l2, MINM, MAXM};
        }
    }
    return np;
}

// ------------------------------------------------------------------------------------------------
// filter kernel: thread per unit, multiword bit-planes in registers
// ------------------------------------------------------------------------------------------------

template <int NW>
__device__ __forceinline__ void load_bits(const u32* __restrict__ plane, u32 pos, u32 (&out)[NW]) {
    const u32* p = plane + (pos >> 5);
    u32 sh = pos & 31;
    u32 prev = __ldg(p);
#pragma unroll
    for (int j = 0; j < NW; j++) {
        u32 nxt = __ldg(p + j + 1);
        out[j] = __funnelshift_r(prev, nxt, sh);
        prev = nxt;
    }
}

template <int NW>
__device__ __forceinline__ void mask_bits(u32 (&x)[NW], int wl) {
#pragma unroll
    for (int j = 0; j < NW; j++) {
        int v = wl - 32 * j;
        x[j] &= v <= 0 ? 0u : low_mask(v);
    }
}

// exclusive prefix XOR over the multiword bit vector: out bit i = xor of in bits [0, i)
template <int NW>
__device__ __forceinline__ void prefix_xor_excl(const u32 (&in)[NW], u32 (&out)[NW]) {
    u32 carry = 0, prev_top = 0;
#pragma unroll
    for (int j = 0; j < NW; j++) {
        u32 inc = prefix_xor32(in[j]) ^ carry;
        out[j] = (inc << 1) | prev_top;
        prev_top = inc >> 31;
        carry = 0u - prev_top;
    }
}

template <int NW>
__device__ __forceinline__ void shr1(u32 (&x)[NW]) {
#pragma unroll
    for (int j = 0; j < NW; j++) x[j] = __funnelshift_r(x[j], j + 1 < NW ? x[j + 1] : 0u, 1);
}

template <int NW>
__device__ __forceinline__ void shr_var(const u32 (&in)[NW], int k, u32 (&out)[NW]) {
#pragma unroll
    for (int j = 0; j < NW; j++) out[j] = in[j];
    for (int s = k >> 5; s > 0; s--) {
#pragma unroll
        for (int j = 0; j < NW; j++) out[j] = j + 1 < NW ? out[j + 1] : 0u;
    }
    u32 r = k & 31;
#pragma unroll
    for (int j = 0; j < NW; j++) out[j] = __funnelshift_r(out[j], j + 1 < NW ? out[j + 1] : 0u, r);
}

// window-valid mask for period k: bit i set iff bases i..i+k-1 are all valid (doubling)
template <int NW>
__device__ __forceinline__ void sliding_and(u32 (&a)[NW], int k) {
    u32 t[NW];
    int w = 1;
    while (2 * w <= k) {
        shr_var<NW>(a, w, t);
#pragma unroll
        for (int j = 0; j < NW; j++) a[j] &= t[j];
        w *= 2;
    }
    if (k > w) {
        shr_var<NW>(a, k - w, t);
#pragma unroll
        for (int j = 0; j < NW; j++) a[j] &= t[j];
    }
}

__device__ __forceinline__ void csa(u32 a, u32 b, u32 c, u32& s, u32& cy) {
    s = a ^ b ^ c;
    cy = (a & b) | (c & (a ^ b));
}

// popcount of a multiword vector with carry-save compression (POPC is quarter-rate; LOP3 is not)
template <int NW>
__device__ __forceinline__ int popc_multi(const u32 (&x)[NW]) {
    if constexpr (NW == 1) return __popc(x[0]);
    else if constexpr (NW == 2) return __popc(x[0]) + __popc(x[1]);
    else if constexpr (NW == 3) {
        u32 s, c; csa(x[0], x[1], x[2], s, c);
        return __popc(s) + 2 * __popc(c);
    } else if constexpr (NW == 5) {
        u32 s1, c1, s2, c2; csa(x[0], x[1], x[2], s1, c1); csa(s1, x[3], x[4], s2, c2);
        return __popc(s2) + 2 * (__popc(c1) + __popc(c2));
    } else if constexpr (NW == 8) {
        u32 s1, c1, s2, c2, s3, c3, s4, c4;
        csa(x[0], x[1], x[2], s1, c1); csa(x[3], x[4], x[5], s2, c2); csa(s1, s2, x[6], s3, c3);
        csa(c1, c2, c3, s4, c4);
        return __popc(s3) + __popc(x[7]) + 2 * __popc(s4) + 4 * __popc(c4);
    } else {
        int t = 0;
#pragma unroll
        for (int j = 0; j < NW; j++) t += __popc(x[j]);
        return t;
    }
}

// Returns true iff some k in [k0, k1] MAY reach the LOW threshold in window [pos, pos + wl).
template <int NW>
__device__ __noinline__ bool probe_filter(const DevBatch& b, u32 pos, int wl, int k0, int k1,
                                          const unsigned short* __restrict__ thr) {
    u32 PH[NW], PL[NW], QH[NW], QL[NW], WV[NW];
    {
        u32 t[NW];
        load_bits<NW>(b.hi, pos, t); prefix_xor_excl<NW>(t, PH);
        load_bits<NW>(b.lo, pos, t); prefix_xor_excl<NW>(t, PL);
        load_bits<NW>(b.val, pos, WV); mask_bits<NW>(WV, wl);
    }
    shr_var<NW>(PH, k0, QH);
    shr_var<NW>(PL, k0, QL);
    sliding_and<NW>(WV, k0);
    for (int k = k0; k <= k1; k++) {
        u32 a[NW], bb[NW], c[NW];
#pragma unroll
        for (int j = 0; j < NW; j++) {
            u32 dh = QH[j] ^ PH[j], dl = QL[j] ^ PL[j];
            a[j] = dh & WV[j]; bb[j] = dl & WV[j]; c[j] = a[j] & dl;
        }
        int T = popc_multi<NW>(WV);
        int cH = popc_multi<NW>(a), cL = popc_multi<NW>(bb), c11 = popc_multi<NW>(c);
        int c10 = cH - c11, c01 = cL - c11, c00 = T - cH - cL + c11;
        int U = max(max(c00, c01), max(c10, c11));
        int need = thr[T];
        if (U >= need) {
            // second level: add the parity of the A count (rare: ~2.5e-4 per (window, k) on random reads)
            u32 A[NW], PA[NW], QA[NW], t[NW];
            load_bits<NW>(b.hi, pos, A); load_bits<NW>(b.lo, pos, t);
#pragma unroll
            for (int j = 0; j < NW; j++) A[j] &= t[j];
            prefix_xor_excl<NW>(A, PA);
            shr_var<NW>(PA, k, QA);
            u32 x11[NW], x10[NW], x01[NW], x00[NW];
#pragma unroll
            for (int j = 0; j < NW; j++) {
                u32 da = QA[j] ^ PA[j];
                u32 dh = QH[j] ^ PH[j], dl = QL[j] ^ PL[j];
                x11[j] = c[j] & da; x10[j] = a[j] & ~dl & da; x01[j] = bb[j] & ~dh & da; x00[j] = WV[j] & ~dh & ~dl & da;
            }
            int n11 = popc_multi<NW>(x11), n10 = popc_multi<NW>(x10), n01 = popc_multi<NW>(x01), n00 = popc_multi<NW>(x00);
            int U2 = max(max(max(n11, c11 - n11), max(n10, c10 - n10)), max(max(n01, c01 - n01), max(n00, c00 - n00)));
            if (U2 >= need) return true;
        }
        shr1<NW>(QH); shr1<NW>(QL);
        u32 t[NW];
#pragma unroll
        for (int j = 0; j < NW; j++) t[j] = WV[j];
        shr1<NW>(t);
#pragma unroll
        for (int j = 0; j < NW; j++) WV[j] &= t[j];
    }
    return false;
}

template <int MAXNW>
__device__ __forceinline__ bool probe_dispatch(const DevBatch& b, const Probe& p, const unsigned short* thr) {
    if (p.k1 < p.k0) return false;
    int need_bits = p.wl + 1;  // prefix planes hold wl + 1 entries
    if (need_bits <= 96) return probe_filter<3>(b, p.pos, p.wl, p.k0, p.k1, thr);
    if constexpr (MAXNW >= 5) { if (need_bits <= 160) return probe_filter<5>(b, p.pos, p.wl, p.k0, p.k1, thr); }
    if constexpr (MAXNW >= 8) { if (need_bits <= 256) return probe_filter<8>(b, p.pos, p.wl, p.k0, p.k1, thr); }
    return true;  // window too long for the bit-parallel filter: let the exact kernel decide
}

template <int MAXNW>
__global__ void __launch_bounds__(256) trew_filter_kernel(DevCfg cfg, DevBatch b, u32 n_units,
                                                          u32* __restrict__ survivors, u32* __restrict__ n_survivors) {
    __shared__ unsigned short thr[kThrTableSize];
    for (int i = threadIdx.x; i < kThrTableSize; i += blockDim.x) thr[i] = cfg.thr_low[i];
    __syncthreads();
    u32 stride = gridDim.x * blockDim.x;
    u32 n_round = (n_units + 31u) & ~31u;
    for (u32 u = blockIdx.x * blockDim.x + threadIdx.x; u < n_round; u += stride) {
        bool maybe = false;
        if (u < n_units) {
            Probe p[4];
            int np = unit_probes(cfg, b, u, p);
            for (int i = 0; i < np && !maybe; i++) maybe = probe_dispatch<MAXNW>(b, p[i], thr);
        }
        u32 m = __ballot_sync(0xffffffffu, maybe);
        if (m) {
            u32 base = 0;
            if (lane_id() == 0) base = atomicAdd(n_survivors, (u32)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (maybe) survivors[base + __popc(m & ((1u << lane_id()) - 1u))] = u;
        }
    }
}

void launch_filter(const DevCfg& cfg, const DevBatch& b, unsigned int n_units, unsigned int max_read_len,
                   unsigned int* survivors, unsigned int* n_survivors, int sm_count, cudaStream_t stream) {
    if (n_units == 0) return;
    // longest probe window: a half read, a whole read (n < 4*MAX) or a slice (long mode)
    unsigned int longest;
    if (cfg.mode == 2) longest = 2u * (unsigned)cfg.slice_len;
    else longest = (max_read_len < 4u * (unsigned)cfg.max_mer) ? max_read_len : (max_read_len + 1) / 2;
    int blocks = sm_count * 8;
    unsigned int need = (n_units + 255) / 256;
    if ((unsigned)blocks > need) blocks = (int)need;
    if (longest + 1 <= 96) trew_filter_kernel<3><<<blocks, 256, 0, stream>>>(cfg, b, n_units, survivors, n_survivors);
    else if (longest + 1 <= 160) trew_filter_kernel<5><<<blocks, 256, 0, stream>>>(cfg, b, n_units, survivors, n_survivors);
    else trew_filter_kernel<8><<<blocks, 256, 0, stream>>>(cfg, b, n_units, survivors, n_survivors);
}

// ------------------------------------------------------------------------------------------------
// exact kernel: warp per survivor unit
// ------------------------------------------------------------------------------------------------

constexpr int kPlaneWords = 36;  // 32 window words + zero padding for shifted reads
#ifndef TREW_EXACT_BPS
#define TREW_EXACT_BPS 8   // resident exact-kernel blocks per SM (latency-bound: occupancy beats registers)
#endif
constexpr int kExactWarps = 4;

struct WarpMem {
    u32* H; u32* L; u32* V; u32* PH; u32* PL;  // kPlaneWords each; word j owned by lane j
    u64* rev2;                                   // kPlaneWords words: reversed, interleaved 2-bit stream
    unsigned short* run_start;                   // [cap]
    unsigned short* run_cw;                      // [cap + 1] ordinal of the run's first window among valid windows
    unsigned short* run_total;                   // [cap] class total for leaders, 0 otherwise
    u64* run_lo; u64* run_hi;                    // [cap] canonical rotation of the run's class
    u32* run_hash;                               // [cap + 4] 32-bit fold of the canonical rotation
    int cap;
};

__host__ __device__ inline size_t exact_warp_bytes(int cap) {
    size_t b = 5 * kPlaneWords * sizeof(u32) + kPlaneWords * sizeof(u64);
    b += (size_t)(3 * cap + 4) * sizeof(unsigned short);
    b = (b + 15) & ~(size_t)15;
    b += (size_t)2 * cap * sizeof(u64);
    b += (size_t)(cap + 4) * sizeof(u32);
    return (b + 15) & ~(size_t)15;
}

size_t exact_smem_bytes(int run_cap, bool) { return exact_warp_bytes(run_cap) * kExactWarps; }

struct KStat { int T, M; u64 s_lo, s_hi; bool homo; int nruns; };

struct ScanRes { int th, tl; u64 sh_lo, sh_hi, sl_lo, sl_hi; };

// ---- k-mer arithmetic (src/kmer.cpp:39-74, 1815-1867) ------------------------------------------

__device__ __noinline__ u64 canon64(u64 w, int k) {
    u64 best = w, cur = w;
    int sh = 2 * (k - 1);
    for (int r = 1; r < k; r++) {
        cur = ((cur & 3ULL) << sh) | (cur >> 2);
        best = cur < best ? cur : best;
    }
    return best;
}

__device__ __noinline__ u128 canon128(u128 w, int k) {
    u128 best = w, cur = w;
    int sh = 2 * (k - 1);
    for (int r = 1; r < k; r++) {
        cur = ((cur & 3) << sh) | (cur >> 2);
        best = cur < best ? cur : best;
    }
    return best;
}

__device__ __forceinline__ u64 rev_pairs64(u64 x) {  // reverse the order of the 32 two-bit symbols
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
}

__device__ __forceinline__ void canon_pair(u64& lo, u64& hi, int k) {
    if (k <= 32) { lo = canon64(lo, k); hi = 0; }
    else { u128 c = canon128(((u128)hi << 64) | lo, k); lo = (u64)c; hi = (u64)(c >> 64); }
}

// canonical rotation of the reverse complement (rot_reverse_complement, src/kmer.cpp:72-74)
__device__ __forceinline__ void crc_pair(u64& lo, u64& hi, int k) {
    if (k <= 32) {
        u64 r = ~rev_pairs64(lo) >> (64 - 2 * k);
        lo = canon64(r, k); hi = 0;
    } else {
        u128 r = ((u128)(~rev_pairs64(lo)) << 64) | (u128)(~rev_pairs64(hi));
        r >>= (128 - 2 * k);
        u128 c = canon128(r, k); lo = (u64)c; hi = (u64)(c >> 64);
    }
}

__device__ __forceinline__ bool homo_pair(u64 lo, u64 hi, int k) {  // get_repeat_check: <= 1 distinct base
    if (k <= 1) return true;
    u128 w = ((u128)hi << 64) | lo;
    u128 m = (((u128)1 << (2 * (k - 1))) - 1);
    return ((w ^ (w >> 2)) & m) == 0;
}

__device__ __forceinline__ bool less_pair(u64 alo, u64 ahi, u64 blo, u64 bhi) {
    return ahi < bhi || (ahi == bhi && alo < blo);
}

// ---- warp state ----------------------------------------------------------------------------------

struct Warp {
    WarpMem m;
    u32 lane;
    u32 cur_pos; int cur_len;  // window currently loaded (cur_len < 0: none)
    u32 h, l, v;               // this lane's word of the window planes
    u32 ev_pos; int ev_len, ev_k;  // (window, k) whose run list is in shared memory (ev_k < 0: none)
    KStat ev;
};

__device__ __forceinline__ u32 shfl_next_bit0(u32 x, u32 lane) {  // bit 0 of the next lane's word (0 for lane 31)
    u32 nx = __shfl_down_sync(0xffffffffu, x, 1);
    return lane == 31 ? 0u : (nx & 1u);
}

// load window [pos, pos+len) of the batch planes; builds prefix-XOR planes and the reversed 2-bit stream
__device__ __noinline__ void load_window(Warp& w, const DevBatch& b, u32 pos, int len) {
    if (w.cur_len == len && w.cur_pos == pos) return;
    __syncwarp();
    const u32 lane = w.lane;
    int vbits = len - 32 * (int)lane;
    u32 msk = vbits <= 0 ? 0u : low_mask(vbits);
    u32 h = 0, l = 0, v = 0;
    if (vbits > 0) {
        u32 wi = (pos >> 5) + lane, sh = pos & 31;
        h = __funnelshift_r(__ldg(b.hi + wi), __ldg(b.hi + wi + 1), sh) & msk;
        l = __funnelshift_r(__ldg(b.lo + wi), __ldg(b.lo + wi + 1), sh) & msk;
        v = __funnelshift_r(__ldg(b.val + wi), __ldg(b.val + wi + 1), sh) & msk;
    }
    w.h = h; w.l = l; w.v = v;
    // exclusive prefix-XOR planes, word j in lane j (len + 1 <= 1024 entries)
    u32 ih = prefix_xor32(h), il = prefix_xor32(l);
    u32 bh = __ballot_sync(0xffffffffu, ih >> 31), bl = __ballot_sync(0xffffffffu, il >> 31);
    u32 lt = (1u << lane) - 1u;
    if (__popc(bh & lt) & 1) ih = ~ih;
    if (__popc(bl & lt) & 1) il = ~il;
    u32 ph_prev = __shfl_up_sync(0xffffffffu, ih, 1), pl_prev = __shfl_up_sync(0xffffffffu, il, 1);
    u32 ph = (ih << 1) | (lane ? ph_prev >> 31 : 0u);
    u32 pl = (il << 1) | (lane ? pl_prev >> 31 : 0u);
    w.m.H[lane] = h; w.m.L[lane] = l; w.m.V[lane] = v; w.m.PH[lane] = ph; w.m.PL[lane] = pl;
    if (lane < kPlaneWords - 32) {
        w.m.H[32 + lane] = 0; w.m.L[32 + lane] = 0; w.m.V[32 + lane] = 0; w.m.PH[32 + lane] = 0; w.m.PL[32 + lane] = 0;
        w.m.rev2[32 + lane] = 0;
    }
    __syncwarp();
    // reversed interleaved stream: 64-bit word j covers reversed positions [32j, 32j+32), i.e. original
    // bases cs .. cs+31 with cs = len - 32j - 32, highest base first (so a k-mer read from it has its
    // first base in the most significant bits, like the reference's shift-in order, src/kmer.cpp:2186-2189)
    {
        int cs = len - 32 * (int)lane - 32;
        u32 hb = 0, lb = 0;
        if (cs > -32) {
            if (cs >= 0) {
                int wi = cs >> 5, sh = cs & 31;
                hb = __funnelshift_r(w.m.H[wi], w.m.H[wi + 1], sh);
                lb = __funnelshift_r(w.m.L[wi], w.m.L[wi + 1], sh);
            } else {
                hb = w.m.H[0] << (-cs);
                lb = w.m.L[0] << (-cs);
            }
        }
        hb = __brev(hb); lb = __brev(lb);
        u64 x = hb, y = lb;
        x = (x | (x << 16)) & 0x0000FFFF0000FFFFULL; y = (y | (y << 16)) & 0x0000FFFF0000FFFFULL;
        x = (x | (x << 8)) & 0x00FF00FF00FF00FFULL;  y = (y | (y << 8)) & 0x00FF00FF00FF00FFULL;
        x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0FULL;  y = (y | (y << 4)) & 0x0F0F0F0F0F0F0F0FULL;
        x = (x | (x << 2)) & 0x3333333333333333ULL;  y = (y | (y << 2)) & 0x3333333333333333ULL;
        x = (x | (x << 1)) & 0x5555555555555555ULL;  y = (y | (y << 1)) & 0x5555555555555555ULL;
        w.m.rev2[lane] = (x << 1) | y;
    }
    __syncwarp();
    w.cur_pos = pos; w.cur_len = len;
}

// k-mer starting at base i (first base most significant)
__device__ __forceinline__ void kmer_at(const Warp& w, int i, int k, u64& lo, u64& hi) {
    int o = 2 * (w.cur_len - i - k);
    int wi = o >> 6, sh = o & 63;
    u64 a = w.m.rev2[wi], b = w.m.rev2[wi + 1];
    lo = sh ? (a >> sh) | (b << (64 - sh)) : a;
    if (k <= 32) {
        if (k < 32) lo &= (1ULL << (2 * k)) - 1ULL;
        hi = 0;
    } else {
        u64 c = w.m.rev2[wi + 2];
        hi = sh ? (b >> sh) | (c << (64 - sh)) : b;
        if (k < 64) hi &= (1ULL << (2 * k - 64)) - 1ULL;
    }
}

// window-valid word for period k from scratch: WV_1 = V, WV_{t+1} = WV_t & (WV_t >> 1)
__device__ __forceinline__ u32 wv_for_k(const Warp& w, int k) {
    u32 wv = w.v;
    for (int t = 1; t < k; t++) {
        u32 nb = shfl_next_bit0(wv, w.lane);
        wv &= (wv >> 1) | (nb << 31);
    }
    return wv;
}

__device__ __forceinline__ u32 wv_step(u32 wv, u32 lane) {
    u32 nb = shfl_next_bit0(wv, lane);
    return wv & ((wv >> 1) | (nb << 31));
}

// upper bound on the largest class count for period k from the 4-bucket parity signature
__device__ __forceinline__ int bound_k(const Warp& w, int k, u32 wv, int T) {
    int s = k >> 5, r = k & 31;
    u32 qh = __funnelshift_r(w.m.PH[w.lane + s], w.m.PH[w.lane + s + 1], r);
    u32 ql = __funnelshift_r(w.m.PL[w.lane + s], w.m.PL[w.lane + s + 1], r);
    u32 dh = (qh ^ w.m.PH[w.lane]) & wv, dl = (ql ^ w.m.PL[w.lane]) & wv;
    u32 packed = (u32)__popc(dh) | ((u32)__popc(dl) << 10) | ((u32)__popc(dh & dl) << 20);
    packed = __reduce_add_sync(0xffffffffu, packed);
    int cH = packed & 1023, cL = (packed >> 10) & 1023, c11 = packed >> 20;
    int c10 = cH - c11, c01 = cL - c11, c00 = T - cH - cL + c11;
    return max(max(c00, c01), max(c10, c11));
}

// Exact class statistics of the loaded window for one period (the inner loops of k_mer_check,
// src/kmer.cpp:2183-2216, without the early break): T valid windows, M largest class, S the class that
// first reaches M.  Leaves the run list in shared memory (run_lo/hi canonical class per run, run_total
// class total on the first run of each class) for emit_classes().
__device__ __noinline__ KStat eval_k(Warp& w, int k, u32 wv) {
    KStat ks; ks.T = 0; ks.M = 0; ks.s_lo = ks.s_hi = 0; ks.homo = false; ks.nruns = 0;
    const u32 lane = w.lane;
    int T = (int)__reduce_add_sync(0xffffffffu, (u32)__popc(wv));
    ks.T = T;
    if (T == 0) return ks;
    // link bits: windows i and i+1 are both valid and base[i] == base[i+k]   (Lemma L1)
    int s = k >> 5, r = k & 31;
    u32 hs = __funnelshift_r(w.m.H[lane + s], w.m.H[lane + s + 1], r);
    u32 ls = __funnelshift_r(w.m.L[lane + s], w.m.L[lane + s + 1], r);
    u32 eq = ~((hs ^ w.h) | (ls ^ w.l));
    u32 nb = shfl_next_bit0(wv, lane);
    u32 link = eq & wv & ((wv >> 1) | (nb << 31));
    u32 link_prev = __shfl_up_sync(0xffffffffu, link, 1);
    u32 rs = wv & ~((link << 1) | (lane ? link_prev >> 31 : 0u));  // run starts
    // exclusive scans over lanes of (valid windows, run starts), packed 16:16
    u32 pk = (u32)__popc(wv) | ((u32)__popc(rs) << 16);
    u32 inc = pk;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, d);
        if ((int)lane >= d) inc += t;
    }
    u32 exc = inc - pk;
    int R = (int)(__shfl_sync(0xffffffffu, inc, 31) >> 16);
    ks.nruns = R;
    {
        u32 cwb = exc & 0xffffu, rb = exc >> 16;
        u32 x = rs;
        while (x) {
            int bit = __ffs(x) - 1;
            x &= x - 1;
            w.m.run_start[rb] = (unsigned short)(32 * lane + bit);
            w.m.run_cw[rb] = (unsigned short)(cwb + __popc(wv & ((1u << bit) - 1u)));
            rb++;
        }
        if (lane == 0) w.m.run_cw[R] = (unsigned short)T;
    }
    __syncwarp();
    // one canonicalisation per run
    for (int q = lane; q < R; q += 32) {
        u64 lo, hi;
        kmer_at(w, w.m.run_start[q], k, lo, hi);
        canon_pair(lo, hi, k);
        w.m.run_lo[q] = lo; w.m.run_hi[q] = hi;
        w.m.run_hash[q] = (u32)lo ^ (u32)(lo >> 32) ^ (u32)hi ^ (u32)(hi >> 32);
    }
    __syncwarp();
    // merge runs of the same class: total windows, ordinal of the class's last window, leader = first run.
    // All pairs, but the inner loop only touches a 32-bit hash per run (4 per 128-bit load); the full key and
    // the counts are read on a hash hit only.
    u32 best = 0; int best_q = -1;
    const bool wide = k > 32;
    const int R4 = (R + 3) & ~3;
    for (int q0 = 0; q0 < R; q0 += 32) {
        int q = q0 + lane;
        if (q < R) {
            u64 mlo = w.m.run_lo[q], mhi = w.m.run_hi[q];
            u32 mh = w.m.run_hash[q];
            int total = 0, last = 0; bool leader = true;
            for (int p4 = 0; p4 < R4; p4 += 4) {
                uint4 hv = *reinterpret_cast<const uint4*>(w.m.run_hash + p4);
                u32 hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    int p = p4 + t;
                    if (hh[t] == mh && p < R && w.m.run_lo[p] == mlo && (!wide || w.m.run_hi[p] == mhi)) {
                        int c0 = w.m.run_cw[p], c1 = w.m.run_cw[p + 1];
                        total += c1 - c0; last = max(last, c1 - 1);
                        if (p < q) leader = false;
                    }
                }
            }
            w.m.run_total[q] = leader ? (unsigned short)total : (unsigned short)0;
            if (leader) {
                // K_MER_DATA_MAX_SEQ: the class whose running count first reaches the final maximum
                // (strict '<' at src/kmer.cpp:2202) = max total, ties broken by the EARLIEST last window
                u32 score = ((u32)total << 10) | (u32)(1023 - last);
                if (score > best) { best = score; best_q = q; }
            }
        }
    }
    u32 wbest = __reduce_max_sync(0xffffffffu, best);
    u32 who = __ballot_sync(0xffffffffu, best == wbest && best_q >= 0);
    int src = __ffs(who) - 1;
    int bq = __shfl_sync(0xffffffffu, best_q, src);
    __syncwarp();
    ks.M = (int)(wbest >> 10);
    ks.s_lo = w.m.run_lo[bq]; ks.s_hi = w.m.run_hi[bq];
    ks.homo = homo_pair(ks.s_lo, ks.s_hi, k);
    return ks;
}

// add every distinct class of the last eval_k() to a result table (optionally RC-folded)
__device__ __noinline__ void emit_classes(const DevCfg& cfg, Warp& w, int k, int nruns, int table, bool folded) {
    u32 meta = ((u32)table << 8) | (u32)k;
    for (int q = w.lane; q < nruns; q += 32) {
        int total = w.m.run_total[q];
        if (total == 0) continue;
        u64 lo = w.m.run_lo[q], hi = w.m.run_hi[q];
        if (folded) {
            u64 rlo = lo, rhi = hi;
            crc_pair(rlo, rhi, k);
            if (less_pair(rlo, rhi, lo, hi)) { lo = rlo; hi = rhi; }
        }
        table_add(cfg, meta, lo, hi, (u64)total);
    }
    __syncwarp();
}

// bit j set for every multiple j of k, j <= 64 (bit 64 does not exist: k = 64 is never a proper divisor target)
__device__ __forceinline__ u64 multiples_mask(int k) {
    u64 m = 0;
    for (int j = k; j < 64; j += k) m |= 1ULL << j;
    return m;
}

// k_mer_check / k_mer_check_128 (src/kmer.cpp:2144-2547) without emission: target_k_high / target_k_low and
// the K_MER_DATA_MAX_SEQ of each.  Periods that cannot be accepted by either selection (divisor rule,
// or the signature bound below the running threshold) are skipped without an exact count.
//
// Pre-test soundness: the reference accepts k iff fl(M/T) >= need.  fl(M/T) >= need implies
// M >= need*T*(1 - 2^-53), and U >= M, so "U >= need*T*(1 - 1e-12)" (evaluated in double, relative error
// ~2^-52) never rejects a period the reference would accept.  The exact test after eval_k uses the same
// IEEE division as the reference.
struct SelState {
    ScanRes res;
    u64 blockedL, blockedH;   // periods with an accepted divisor (k % tk == 0, src/kmer.cpp:2225-2230)
    bool blk64L, blk64H;
    double needL, needH;      // max(baseline, last accepted frequency)
};

// one exact evaluation + the two acceptance tests of src/kmer.cpp:2221-2258
__device__ __noinline__ void consider(Warp& w, SelState& st, u32 pos, int len, int k, u32 wv) {
    KStat ks = eval_k(w, k, wv);
    w.ev = ks; w.ev_pos = pos; w.ev_len = len; w.ev_k = k;
    if (ks.homo) return;
    bool blkL = k < 64 ? ((st.blockedL >> k) & 1ULL) != 0 : st.blk64L;
    bool blkH = k < 64 ? ((st.blockedH >> k) & 1ULL) != 0 : st.blk64H;
    double f = (double)ks.M / (double)ks.T;
    bool accL = !blkL && f >= st.needL, accH = !blkH && f >= st.needH;
    if (accL || accH) {
        u64 mm = multiples_mask(k);
        bool m64 = (64 % k) == 0;
        if (accL) { st.res.tl = k; st.needL = f; st.blockedL |= mm; st.blk64L |= m64; st.res.sl_lo = ks.s_lo; st.res.sl_hi = ks.s_hi; }
        if (accH) { st.res.th = k; st.needH = f; st.blockedH |= mm; st.blk64H |= m64; st.res.sh_lo = ks.s_lo; st.res.sh_hi = ks.s_hi; }
    }
}

__device__ __noinline__ ScanRes scan_stats(const DevCfg& cfg, Warp& w, const DevBatch& b, u32 pos, int len, int kmin, int kmax) {
    SelState st;
    st.res.th = st.res.tl = 0; st.res.sh_lo = st.res.sh_hi = st.res.sl_lo = st.res.sl_hi = 0;
    if (kmax < kmin) return st.res;
    load_window(w, b, pos, len);
    st.blockedL = st.blockedH = 0; st.blk64L = st.blk64H = false;
    st.needL = cfg.low; st.needH = cfg.high;
    const double slack = 1.0 - 1e-12;
    const u32 lane = w.lane;

    // Visit, in ascending order, the periods of one 32-wide block whose bound reached the LOW threshold.
    // U / T are per-lane (lane <-> period kb + lane); wv_of(bit) yields this lane's window-valid word.
    auto visit = [&](int kb, u32 cm, int U, int T, auto wv_of) {
        while (cm) {
            int bit = __ffs(cm) - 1;
            cm &= cm - 1;
            int kk = kb + bit;
            bool blkL = kk < 64 ? ((st.blockedL >> kk) & 1ULL) != 0 : st.blk64L;
            bool blkH = kk < 64 ? ((st.blockedH >> kk) & 1ULL) != 0 : st.blk64H;
            int Uk = __shfl_sync(0xffffffffu, U, bit), Tk = __shfl_sync(0xffffffffu, T, bit);
            u32 wv = wv_of(bit, Tk);
            if (blkL && blkH) continue;
            double dU = (double)Uk, dT = (double)Tk;
            bool candL = !blkL && dU >= st.needL * dT * slack, candH = !blkH && dU >= st.needH * dT * slack;
            if (!candL && !candH) continue;
            consider(w, st, pos, len, kk, wv);
        }
    };

    if (len <= 127) {
        // Short window (the 75 / 150-base case): every lane bounds its own period (lane <-> k) with the
        // window's planes held in registers, invalid bases included; no cross-lane traffic until a
        // period qualifies.
        u32 ph[5], pl[5], vv[4];
#pragma unroll
        for (int j = 0; j < 5; j++) { ph[j] = w.m.PH[j]; pl[j] = w.m.PL[j]; }
#pragma unroll
        for (int j = 0; j < 4; j++) vv[j] = w.m.V[j];
        for (int kb = kmin; kb <= kmax; kb += 32) {
            const int k = kb + (int)lane;
            u32 wvv[4] = {0, 0, 0, 0};
            int U = 0, T = 0;
            if (k <= kmax) {
#pragma unroll
                for (int j = 0; j < 4; j++) wvv[j] = vv[j];
                sliding_and<4>(wvv, k);
                u32 qh[5], ql[5];
                shr_var<5>(ph, k, qh);
                shr_var<5>(pl, k, ql);
                int cH = 0, cL = 0, c11 = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    u32 dh = (qh[j] ^ ph[j]) & wvv[j], dl = (ql[j] ^ pl[j]) & wvv[j];
                    T += __popc(wvv[j]); cH += __popc(dh); cL += __popc(dl); c11 += __popc(dh & dl);
                }
                int c10 = cH - c11, c01 = cL - c11, c00 = T - cH - cL + c11;
                U = max(max(c00, c01), max(c10, c11));
            }
            bool cand = T > 0 && (double)U >= cfg.low * (double)T * slack;
            u32 cm = __ballot_sync(0xffffffffu, cand);
            visit(kb, cm, U, T, [&](int bit, int) {
                u32 a0 = __shfl_sync(0xffffffffu, wvv[0], bit), a1 = __shfl_sync(0xffffffffu, wvv[1], bit);
                u32 a2 = __shfl_sync(0xffffffffu, wvv[2], bit), a3 = __shfl_sync(0xffffffffu, wvv[3], bit);
                return lane == 0 ? a0 : lane == 1 ? a1 : lane == 2 ? a2 : lane == 3 ? a3 : 0u;
            });
        }
        return st.res;
    }

    const bool all_valid = (int)__reduce_add_sync(0xffffffffu, (u32)__popc(w.v)) == len;
    if (all_valid) {
        // Long window without invalid bases: the valid windows of period k are positions [0, len - k].
        for (int kb = kmin; kb <= kmax; kb += 32) {
            const int k = kb + (int)lane;
            const int T = (k <= kmax && len - k + 1 > 0) ? len - k + 1 : 0;
            int U = 0;
            if (T > 0) {
                const int s = k >> 5, r = k & 31;
                int cH = 0, cL = 0, c11 = 0;
                for (int j = 0; j * 32 < T; j++) {
                    u32 wvj = low_mask(min(32, T - 32 * j));
                    u32 dh = (__funnelshift_r(w.m.PH[j + s], w.m.PH[j + s + 1], r) ^ w.m.PH[j]) & wvj;
                    u32 dl = (__funnelshift_r(w.m.PL[j + s], w.m.PL[j + s + 1], r) ^ w.m.PL[j]) & wvj;
                    cH += __popc(dh); cL += __popc(dl); c11 += __popc(dh & dl);
                }
                int c10 = cH - c11, c01 = cL - c11, c00 = T - cH - cL + c11;
                U = max(max(c00, c01), max(c10, c11));
            }
            bool cand = T > 0 && (double)U >= cfg.low * (double)T * slack;
            u32 cm = __ballot_sync(0xffffffffu, cand);
            visit(kb, cm, U, T, [&](int, int Tk) {
                int vb = Tk - 32 * (int)lane;
                return vb <= 0 ? 0u : low_mask(min(32, vb));
            });
        }
        return st.res;
    }

    // windows with invalid bases: walk the periods in order, keeping the window-valid mask incrementally
    u32 wv = wv_for_k(w, kmin);
    for (int k = kmin; k <= kmax; k++, wv = wv_step(wv, lane)) {
        bool blkL = k < 64 ? ((st.blockedL >> k) & 1ULL) != 0 : st.blk64L;
        bool blkH = k < 64 ? ((st.blockedH >> k) & 1ULL) != 0 : st.blk64H;
        if (blkL && blkH) continue;
        int T = (int)__reduce_add_sync(0xffffffffu, (u32)__popc(wv));
        if (T == 0) continue;
        int U = bound_k(w, k, wv, T);
        double dU = (double)U, dT = (double)T;
        bool candL = !blkL && dU >= st.needL * dT * slack, candH = !blkH && dU >= st.needH * dT * slack;
        if (!candL && !candH) continue;
        consider(w, st, pos, len, k, wv);
    }
    return st.res;
}

// class statistics + run list of (window, k), re-using the last evaluation when it is the same one
__device__ __noinline__ KStat eval_cached(Warp& w, const DevBatch& b, u32 pos, int len, int k) {
    if (w.ev_k == k && w.ev_pos == pos && w.ev_len == len) return w.ev;
    load_window(w, b, pos, len);
    u32 wv = wv_for_k(w, k);
    w.ev = eval_k(w, k, wv);
    w.ev_pos = pos; w.ev_len = len; w.ev_k = k;
    return w.ev;
}

// the emission half of k_mer_check for one target k (src/kmer.cpp:2264-2328): every class, un-folded unless asked
__device__ void emit_window(const DevCfg& cfg, Warp& w, const DevBatch& b, u32 pos, int len, int k, int table, bool folded) {
    KStat ks = eval_cached(w, b, pos, len, k);
    emit_classes(cfg, w, k, ks.nruns, table, folded);
}

// k_mer_target / k_mer_target_128 (src/kmer.cpp:1894-2142)
__device__ void target_window(const DevCfg& cfg, Warp& w, const DevBatch& b, u32 pos, int len, int k, double B, int table) {
    KStat ks = eval_cached(w, b, pos, len, k);
    if (ks.T > 0 && !ks.homo && (double)ks.M / (double)ks.T >= B) emit_classes(cfg, w, k, ks.nruns, table, true);
}

// ---- routing -----------------------------------------------------------------------------------

enum { T_F = 0, T_B = 2, T_O = 4 };

// buffer_task (src/kmer.cpp:80-266)
__device__ void route_short(const DevCfg& cfg, Warp& w, const DevBatch& b, u32 u) {
    const int MINM = cfg.min_mer, MAXM = cfg.max_mer;
    u32 b0 = __ldg(b.bit_off + u);
    int n = (int)(__ldg(b.bit_off + u + 1) - b0);
    if (n < 2 * MINM || n > kMaxWindow) return;
    int L[2] = {0, 0}, R[2] = {0, 0};
    if (n >= 4 * MINM) {
        int kmax = min(n / 4, MAXM);
        u32 lpos = b0, rpos = b0 + (u32)(n - (n + 1) / 2);
        int llen = n / 2, rlen = (n + 1) / 2;
        ScanRes l = scan_stats(cfg, w, b, lpos, llen, MINM, kmax);
        ScanRes r = scan_stats(cfg, w, b, rpos, rlen, MINM, kmax);  // always evaluated
        L[0] = l.th; L[1] = l.tl; R[0] = r.th; R[1] = r.tl;
        // right-half emissions survive only for classes where the left half found nothing
        // (nullptr maps at src/kmer.cpp:125, result.backward at :158)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (L[c] > 0 && L[c] == R[c]) target_window(cfg, w, b, b0, n, L[c], c == 0 ? cfg.high : cfg.low, T_O + c);
            else if (L[c] > 0) emit_window(cfg, w, b, lpos, llen, L[c], T_F + c, false);
            else if (R[c] > 0) emit_window(cfg, w, b, rpos, rlen, R[c], T_B + c, false);
        }
    }
    bool hc[2] = {L[0] == 0 && R[0] == 0, L[1] == 0 && R[1] == 0};
    if (4 * MAXM > n && (hc[0] || hc[1])) {
        ScanRes s = scan_stats(cfg, w, b, b0, n, max(n / 4 + 1, MINM), min(n / 2, MAXM));
        if (hc[0] && s.th) emit_window(cfg, w, b, b0, n, s.th, T_O + 0, false);  // un-folded into 'both'
        if (hc[1] && s.tl) emit_window(cfg, w, b, b0, n, s.tl, T_O + 1, false);
    }
}

// buffer_task_pair (src/kmer.cpp:268-745); follows the 128-bit path where the two differ (temp map cleared
// after the large-k block, src/kmer.cpp:722-723)
__device__ void route_pair(const DevCfg& cfg, Warp& w, const DevBatch& b, u32 u) {
    const int MINM = cfg.min_mer, MAXM = cfg.max_mer;
    u32 a0 = __ldg(b.bit_off + 2 * u), a1 = __ldg(b.bit_off + 2 * u + 1), a2 = __ldg(b.bit_off + 2 * u + 2);
    int n1 = (int)(a1 - a0), n2 = (int)(a2 - a1);
    int n = min(n1, n2);
    if (n < 2 * MINM || n1 > kMaxWindow || n2 > kMaxWindow) return;
    int lef[2] = {0, 0}, km[2] = {0, 0};
    if (n >= 4 * MINM) {
        u32 spos[5] = {0, a0, a0 + (u32)(n1 - (n1 + 1) / 2), a1 + (u32)(n2 - (n2 + 1) / 2), a1};
        int slen[5] = {0, n1 / 2, (n1 + 1) / 2, (n2 + 1) / 2, n2 / 2};
        int kmax = min(n / 4, MAXM);
        ScanRes sr[5]; bool have[5] = {false, false, false, false, false};
        // pending emissions per class: segment + temp map id (0 = left, 1 = right)
        int pseg[2][8], ptmp[2][8], np[2] = {0, 0};
        int si[2] = {1, 1}; bool ended[2] = {false, false};
        u64 ks_lo[2] = {0, 0}, ks_hi[2] = {0, 0};
        for (int ti = 1; ti <= 4 && !(ended[0] && ended[1]); ti++) {
            if (!have[ti]) { sr[ti] = scan_stats(cfg, w, b, spos[ti], slen[ti], MINM, kmax); have[ti] = true; }
            int k[2] = {sr[ti].th, sr[ti].tl};
            u64 slo[2] = {sr[ti].sh_lo, sr[ti].sl_lo}, shi[2] = {sr[ti].sh_hi, sr[ti].sl_hi};
#pragma unroll
            for (int c = 0; c < 2; c++) {
                if (!ended[c] && k[c]) { pseg[c][np[c]] = ti; ptmp[c][np[c]] = ti <= 2 ? 0 : 1; np[c]++; }  // emission before the test
                bool ok = !ended[c] && k[c] > 0;
                if (ok && ti != 1) {
                    u64 dlo = slo[c], dhi = shi[c];
                    if (ti > 2) crc_pair(dlo, dhi, k[c]);  // get_dir_seq, src/kmer.cpp:307-313
                    ok = km[c] == k[c] && ks_lo[c] == dlo && ks_hi[c] == dhi;
                }
                if (ok) { si[c]++; km[c] = k[c]; if (ti == 1) { ks_lo[c] = slo[c]; ks_hi[c] = shi[c]; } }
                else ended[c] = true;
            }
        }
        lef[0] = km[0]; lef[1] = km[1];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (si[c] == 5) {
                for (int e = 0; e < np[c]; e++) {
                    int sg = pseg[c][e];
                    emit_window(cfg, w, b, spos[sg], slen[sg], c == 0 ? sr[sg].th : sr[sg].tl, T_O + c, true);
                }
            }
        }
        if (si[0] <= 4 || si[1] <= 4) {
            int sj[2] = {4, 4}; km[0] = km[1] = 0; ended[0] = ended[1] = false;
            for (int tj = 4; tj >= 1 && !(ended[0] && ended[1]); tj--) {
                if (!have[tj]) { sr[tj] = scan_stats(cfg, w, b, spos[tj], slen[tj], MINM, kmax); have[tj] = true; }
                int k[2] = {sr[tj].th, sr[tj].tl};
                u64 slo[2] = {sr[tj].sh_lo, sr[tj].sl_lo}, shi[2] = {sr[tj].sh_hi, sr[tj].sl_hi};
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    if (!ended[c] && k[c]) { pseg[c][np[c]] = tj; ptmp[c][np[c]] = tj <= 2 ? 1 : 0; np[c]++; }  // temps swapped
                    bool ok = sj[c] >= si[c] && !ended[c] && k[c] > 0;
                    if (ok && tj != 4) {
                        u64 dlo = slo[c], dhi = shi[c];
                        if (tj <= 2) crc_pair(dlo, dhi, k[c]);
                        ok = km[c] == k[c] && ks_lo[c] == dlo && ks_hi[c] == dhi;
                    }
                    if (ok) { sj[c]--; km[c] = k[c]; if (tj == 4) { ks_lo[c] = slo[c]; ks_hi[c] = shi[c]; } }
                    else ended[c] = true;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (si[c] <= 4) {
                for (int e = 0; e < np[c]; e++) {
                    int sg = pseg[c][e];
                    emit_window(cfg, w, b, spos[sg], slen[sg], c == 0 ? sr[sg].th : sr[sg].tl,
                                (ptmp[c][e] == 0 ? T_F : T_B) + c, false);
                }
            }
        }
    }
    if (4 * MAXM > n && (lef[0] == 0 || lef[1] == 0 || km[0] == 0 || km[1] == 0)) {
        int lo = max(n / 4 + 1, MINM), hi = min(n / 2, MAXM);
        ScanRes l, r; l.th = l.tl = r.th = r.tl = 0; l.sh_lo = l.sh_hi = l.sl_lo = l.sl_hi = 0; r = l;
        if (lef[0] == 0 || lef[1] == 0) l = scan_stats(cfg, w, b, a0, n1, lo, hi);
        if (km[0] == 0 || km[1] == 0) r = scan_stats(cfg, w, b, a1, n2, lo, hi);
        int ltk[2] = {l.th, l.tl}, rtk[2] = {r.th, r.tl};
        u64 llo[2] = {l.sh_lo, l.sl_lo}, lhi[2] = {l.sh_hi, l.sl_hi}, rlo[2] = {r.sh_lo, r.sl_lo}, rhi[2] = {r.sh_hi, r.sl_hi};
#pragma unroll
        for (int c = 0; c < 2; c++) {
            bool el = lef[c] == 0 && ltk[c] > 0, er = km[c] == 0 && rtk[c] > 0;  // both land in the 'left' temp map
            bool both = lef[c] == 0 && km[c] == 0 && ltk[c] == rtk[c] && ltk[c] > 0;
            if (both) {
                u64 dlo = rlo[c], dhi = rhi[c];
                crc_pair(dlo, dhi, rtk[c]);
                both = llo[c] == dlo && lhi[c] == dhi;
            }
            if (both) {
                if (el) emit_window(cfg, w, b, a0, n1, ltk[c], T_O + c, true);
                if (er) emit_window(cfg, w, b, a1, n2, rtk[c], T_O + c, true);
            }
            if (el) emit_window(cfg, w, b, a0, n1, ltk[c], T_F + c, false);
            if (er) emit_window(cfg, w, b, a1, n2, rtk[c], T_F + c, false);
        }
    }
}

// buffer_task_long (src/kmer.cpp:747-985)
__device__ void route_long(const DevCfg& cfg, Warp& w, const DevBatch& b, u32 u, unsigned char* scratch) {
    const int MINM = cfg.min_mer, MAXM = cfg.max_mer, SL = cfg.slice_len;
    u32 b0 = __ldg(b.bit_off + u);
    int n = (int)(__ldg(b.bit_off + u + 1) - b0);
    if (n < SL) return;  // the reader drops these (src/kmer.cpp:1184)
    int snum = n / SL, mid = (snum + 1) / 2, bonus = n % SL;
    auto s_start = [&](int t) { return (u32)((t - 1) * SL + (t > mid ? bonus : 0)); };
    auto s_len = [&](int t) { return SL + (t == mid ? bonus : 0); };
    // forward walk, pass 1: statistics only; the destination of its emissions is known at its end
    int si[2] = {1, 1}, km[2] = {0, 0}; bool ended[2] = {false, false};
    int nf = 0;
    for (int ti = 1; ti <= snum && !(ended[0] && ended[1]); ti++) {
        ScanRes sr = scan_stats(cfg, w, b, b0 + s_start(ti), s_len(ti), MINM, MAXM);
        if (w.lane == 0) { scratch[2 * (ti - 1)] = (unsigned char)sr.th; scratch[2 * (ti - 1) + 1] = (unsigned char)sr.tl; }
        int k[2] = {sr.th, sr.tl};
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (!ended[c] && k[c] > 0 && (ti == 1 || km[c] == k[c])) { si[c]++; km[c] = k[c]; }
            else ended[c] = true;
        }
        nf = ti;
    }
    __syncwarp();
    // pass 2: replay the walk and emit into 'both' (folded) when every slice survived, else 'forward'
    {
        bool en[2] = {false, false}; int kk[2] = {0, 0};
        bool full[2] = {si[0] == snum + 1, si[1] == snum + 1};
        for (int ti = 1; ti <= nf; ti++) {
            int k[2] = {scratch[2 * (ti - 1)], scratch[2 * (ti - 1) + 1]};
#pragma unroll
            for (int c = 0; c < 2; c++) {
                if (!en[c] && k[c]) emit_window(cfg, w, b, b0 + s_start(ti), s_len(ti), k[c], (full[c] ? T_O : T_F) + c, full[c]);
                if (!en[c] && k[c] > 0 && (ti == 1 || kk[c] == k[c])) kk[c] = k[c];
                else en[c] = true;
            }
        }
    }
    if (si[0] <= snum || si[1] <= snum) {
        int sj[2] = {snum, snum}; km[0] = km[1] = 0; ended[0] = ended[1] = false;
        for (int tj = snum; tj >= 1 && !(ended[0] && ended[1]); tj--) {
            ScanRes sr = scan_stats(cfg, w, b, b0 + s_start(tj), s_len(tj), MINM, MAXM);
            int k[2] = {sr.th, sr.tl};
#pragma unroll
            for (int c = 0; c < 2; c++) {
                if (!ended[c] && k[c]) emit_window(cfg, w, b, b0 + s_start(tj), s_len(tj), k[c], T_B + c, false);  // straight into 'backward'
                if (sj[c] >= si[c] && !ended[c] && k[c] > 0 && (tj == snum || km[c] == k[c])) { sj[c]--; km[c] = k[c]; }
                else ended[c] = true;
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(kExactWarps * 32, TREW_EXACT_BPS) trew_exact_kernel(DevCfg cfg, DevBatch b, ExactArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wid = threadIdx.x >> 5;
    unsigned char* base = smem + (size_t)wid * exact_warp_bytes(a.run_cap);
    Warp w;
    w.lane = lane_id();
    w.m.cap = a.run_cap;
    w.m.H = (u32*)base; w.m.L = w.m.H + kPlaneWords; w.m.V = w.m.L + kPlaneWords;
    w.m.PH = w.m.V + kPlaneWords; w.m.PL = w.m.PH + kPlaneWords;
    w.m.rev2 = (u64*)(w.m.PL + kPlaneWords);
    w.m.run_start = (unsigned short*)(w.m.rev2 + kPlaneWords);
    w.m.run_cw = w.m.run_start + a.run_cap;
    w.m.run_total = w.m.run_cw + a.run_cap + 1;
    size_t off = 5 * kPlaneWords * sizeof(u32) + kPlaneWords * sizeof(u64) + (size_t)(3 * a.run_cap + 4) * sizeof(unsigned short);
    off = (off + 15) & ~(size_t)15;
    w.m.run_lo = (u64*)(base + off);
    w.m.run_hi = w.m.run_lo + a.run_cap;
    w.m.run_hash = (u32*)(w.m.run_hi + a.run_cap);
    w.cur_len = -1; w.cur_pos = 0; w.ev_k = -1; w.ev_pos = 0; w.ev_len = -1;
    w.h = w.l = w.v = 0;
    const u32 n = *a.n_survivors;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.total_survivors) atomicAdd(a.total_survivors, (u64)n);
    unsigned char* scratch = a.slice_scratch + (size_t)(blockIdx.x * kExactWarps + wid) * a.slice_scratch_stride;
    for (;;) {
        u32 idx = 0;
        if (w.lane == 0) idx = atomicAdd(a.work_counter, 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= n) break;
        u32 u = a.survivors[idx];
        if constexpr (MODE == 0) route_short(cfg, w, b, u);
        else if constexpr (MODE == 1) route_pair(cfg, w, b, u);
        else route_long(cfg, w, b, u, scratch);
    }
}

constexpr int kExactBlocksPerSM = TREW_EXACT_BPS;
int exact_warps_total(int sm_count) { return sm_count * kExactBlocksPerSM * kExactWarps; }

cudaError_t prepare_exact(int run_cap_max) {
    int bytes = (int)exact_smem_bytes(run_cap_max, true);
    cudaError_t e = cudaFuncSetAttribute(trew_exact_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(trew_exact_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(trew_exact_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e;
}

void launch_exact(const DevCfg& cfg, const DevBatch& b, const ExactArgs& a, int sm_count, cudaStream_t stream) {
    size_t smem = exact_smem_bytes(a.run_cap, true);
    dim3 grid(sm_count * kExactBlocksPerSM), block(kExactWarps * 32);
    if (cfg.mode == 0) trew_exact_kernel<0><<<grid, block, smem, stream>>>(cfg, b, a);
    else if (cfg.mode == 1) trew_exact_kernel<1><<<grid, block, smem, stream>>>(cfg, b, a);
    else trew_exact_kernel<2><<<grid, block, smem, stream>>>(cfg, b, a);
}

}  // namespace trew
