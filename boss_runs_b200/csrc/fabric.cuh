// fabric.cuh — the strategy update's exchange steps done by the GPUs themselves over peer memory.
//
// The sharded update (include/bossgpu.h, bossgpu_update_fused_*) has four points where shards need each
// other's results; upstream has none of them because it is one process (boss/runs/core.py:160-198 runs the
// per-contig loops and the global threshold back to back). Payloads are tiny (4 B ... 26 KB), so what matters
// is latency: every shard owns one exchange block in its HBM that its peers map (CUDA IPC between processes
// on the NVSwitch box, plain pointers between the virtual shards of one process). A step is ONE small kernel:
//
//   push   store my contribution into every peer's block (remote stores over NVLink), slot [epoch parity][me]
//   signal fence.sys, then store the epoch into every peer's flag word [step][me]
//   wait   spin (ld.acquire.sys on my own block) until every peer's flag reached the epoch
//   reduce combine the slots in shard order into this shard's own state
//
//   step 0  bucket switch flag: max over shards (core.py:110-111 looks at ALL contigs); bin halos of split
//           contigs to the two neighbours (support of the box sums, reference.py:233-260)
//   step 1  normaliser = max benefit (sequences.py:588): max of the double's bit pattern
//   step 2  exponent histogram + ubar0 limbs (sequences.py:596-629): integer sums, exact and order-free
//   step 3  "my packed mask is complete" — k_distribute then reads the few merged rows that live on a
//           neighbour straight from that neighbour's block (Q2 row shift, core.py:141,155)
//
// No host round trip and no library collective sits inside an update: the host enqueues all kernels of the
// update at once and synchronises once at the end.
//
// Reuse safety: small slots are double-buffered by epoch parity. A shard can only reach epoch k+1's push after
// its own wait of epoch k, i.e. after every peer has *pushed* epoch k, which each peer does after it finished
// *reading* epoch k-1 (stream order) — the slot of parity (k+1)&1 is free by then. The mask is single-buffered:
// a shard rewrites it in epoch k+1 only after step 0 of k+1 completed, which every peer enters after its own
// k_distribute of epoch k (stream order again).
//
// A spin that sees no progress for `timeout_ns` gives up, records BOSSGPU_EPEER in UpdateDev::fabric_err and
// lets the update drain; the host reports the error instead of hanging the GPU.
#pragma once
#include "common.cuh"

namespace boss {

constexpr int FAB_MAX_SHARDS = 64;
constexpr int FAB_STEPS = 4;
constexpr int FAB_THREADS = 512;
constexpr int HIST_WORDS = 3 * HBINS + 4;

// byte offsets inside every shard's exchange block; identical on all shards (same nb, halo, n_shards, stride)
struct FabricLayout {
    size_t o_flag;     // u32 [FAB_STEPS][FAB_MAX_SHARDS]      written by peers
    size_t o_sw;       // i32 [2][FAB_MAX_SHARDS]
    size_t o_norm;     // u64 [2][FAB_MAX_SHARDS]
    size_t o_hist;     // u64 [2][n_shards][HIST_WORDS]
    size_t o_halo;     // f64 [2][2 = from left, from right][halo_bins * nb]
    size_t o_mask;     // u8  [mask_stride]   own packed mask, read by peers
    size_t bytes;
};

inline FabricLayout fabric_layout(int n_shards, int nb, int halo_bins, int64_t mask_stride) {
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    FabricLayout L;
    L.o_flag = 0;
    L.o_sw = up(L.o_flag + sizeof(uint32_t) * FAB_STEPS * FAB_MAX_SHARDS);
    L.o_norm = up(L.o_sw + sizeof(int32_t) * 2 * FAB_MAX_SHARDS);
    L.o_hist = up(L.o_norm + sizeof(unsigned long long) * 2 * FAB_MAX_SHARDS);
    L.o_halo = up(L.o_hist + sizeof(unsigned long long) * 2 * (size_t)n_shards * HIST_WORDS);
    L.o_mask = up(L.o_halo + sizeof(double) * 4 * (size_t)(halo_bins > 0 ? halo_bins : 1) * nb);
    L.bytes = up(L.o_mask + (size_t)mask_stride);
    return L;
}

struct FabricArgs {
    char* const* peer;            // device array [n]: every shard's block as THIS device addresses it (peer[me] = own)
    FabricLayout L;
    int n, me;
    unsigned epoch;               // update counter, identical on every shard, starts at 1
    unsigned long long timeout_ns;
    UpdateDev* upd;
    // step 0
    double* ds;                   // [nb][ds_len]
    int64_t ds_len;
    int nb, halo_bins;
    int64_t first_ds_off, first_n_bins;   // first segment (left edge of the shard)
    int64_t last_ds_off, last_n_bins;     // last segment (right edge)
    int send_left, send_right;    // the edge cuts a contig: the neighbour needs my edge bins
    // step 2
    unsigned long long* hist;     // [HIST_WORDS]
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <typename T>
__device__ __forceinline__ T ld_cg(const T* p) { return __ldcg(p); }

// signal every peer, then wait for every peer (threads 0..n-1 own one peer each)
__device__ __forceinline__ void fabric_signal_and_wait(const FabricArgs& a, int step) {
    __syncthreads();                                   // all pushes of this CTA are issued
    const int r = threadIdx.x;
    if (r < a.n && r != a.me) {
        __threadfence_system();                        // ... and ordered before the flag, system-wide
        unsigned* theirs = reinterpret_cast<unsigned*>(a.peer[r] + a.L.o_flag) + step * FAB_MAX_SHARDS + a.me;
        st_release_sys(theirs, a.epoch);
        const unsigned* mine = reinterpret_cast<const unsigned*>(a.peer[a.me] + a.L.o_flag) + step * FAB_MAX_SHARDS + r;
        const unsigned long long t0 = global_ns();
        unsigned spins = 0;
        // epochs only grow; a peer can be at most one update ahead
        while ((int)(ld_acquire_sys(mine) - a.epoch) < 0) {
            if ((++spins & 1023u) == 0 && global_ns() - t0 > a.timeout_ns) {
                atomicExch(&a.upd->fabric_err, BOSSGPU_EPEER);
                break;
            }
            __nanosleep(40);
        }
    }
    __syncthreads();                                   // the acquiring threads hand the data to the whole CTA
}

// ---- step 0: switch flag + bin halos ----------------------------------------------------------------------
__global__ void __launch_bounds__(FAB_THREADS)
k_fabric_switch_halo(FabricArgs a) {
    const int par = a.epoch & 1;
    const int t = threadIdx.x;
    if (t < a.n)
        reinterpret_cast<int32_t*>(a.peer[t] + a.L.o_sw)[par * FAB_MAX_SHARDS + a.me] = a.upd->switched_on;
    const int hb = a.halo_bins;
    const size_t side = (size_t)(hb > 0 ? hb : 1) * a.nb;
    if (hb > 0) {
        // my first bins are the right halo of the left neighbour; my last bins (right-aligned) the left halo
        // of the right neighbour. Bins the segment does not have arrive as 0 (the box sums see nothing there).
        if (a.send_left && a.me > 0) {
            double* dst = reinterpret_cast<double*>(a.peer[a.me - 1] + a.L.o_halo) + ((size_t)par * 2 + 1) * side;
            const int64_t n = a.first_n_bins < hb ? a.first_n_bins : hb;
            for (int i = t; i < hb * a.nb; i += FAB_THREADS) {
                const int b = i / hb, j = i - b * hb;
                dst[i] = j < n ? a.ds[(size_t)b * a.ds_len + a.first_ds_off + j] : 0.0;
            }
        }
        if (a.send_right && a.me + 1 < a.n) {
            double* dst = reinterpret_cast<double*>(a.peer[a.me + 1] + a.L.o_halo) + ((size_t)par * 2 + 0) * side;
            const int64_t n = a.last_n_bins < hb ? a.last_n_bins : hb;
            for (int i = t; i < hb * a.nb; i += FAB_THREADS) {
                const int b = i / hb, j = i - b * hb;                  // slot j holds bin n_bins - hb + j
                dst[i] = j >= hb - n ? a.ds[(size_t)b * a.ds_len + a.last_ds_off + a.last_n_bins - hb + j] : 0.0;
            }
        }
    }
    fabric_signal_and_wait(a, 0);
    if (t == 0) {
        const int32_t* sw = reinterpret_cast<const int32_t*>(a.peer[a.me] + a.L.o_sw) + par * FAB_MAX_SHARDS;
        int32_t on = 0;
        for (int r = 0; r < a.n; ++r) on |= ld_cg(sw + r) != 0;
        a.upd->switched_on = on;
    }
    if (hb > 0) {
        const double* from_l = reinterpret_cast<const double*>(a.peer[a.me] + a.L.o_halo) + ((size_t)par * 2 + 0) * side;
        const double* from_r = from_l + side;
        // the left neighbour sends iff its right edge cuts a contig, i.e. iff my left edge does (and vice versa)
        if (a.send_left && a.me > 0)
            for (int i = t; i < hb * a.nb; i += FAB_THREADS) {
                const int b = i / hb, j = i - b * hb;
                a.ds[(size_t)b * a.ds_len + a.first_ds_off - hb + j] = ld_cg(from_l + i);
            }
        if (a.send_right && a.me + 1 < a.n)
            for (int i = t; i < hb * a.nb; i += FAB_THREADS) {
                const int b = i / hb, j = i - b * hb;
                a.ds[(size_t)b * a.ds_len + a.last_ds_off + a.last_n_bins + j] = ld_cg(from_r + i);
            }
    }
}

// ---- step 1: normaliser -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(FAB_THREADS)
k_fabric_norm(FabricArgs a) {
    const int par = a.epoch & 1;
    const int t = threadIdx.x;
    if (t < a.n)
        reinterpret_cast<unsigned long long*>(a.peer[t] + a.L.o_norm)[par * FAB_MAX_SHARDS + a.me] = a.upd->norm_bits;
    fabric_signal_and_wait(a, 1);
    if (t == 0) {
        const unsigned long long* nm = reinterpret_cast<const unsigned long long*>(a.peer[a.me] + a.L.o_norm) + par * FAB_MAX_SHARDS;
        unsigned long long mx = 0ull;                  // non-negative doubles order like their bit patterns
        for (int r = 0; r < a.n; ++r) { const unsigned long long v = ld_cg(nm + r); mx = v > mx ? v : mx; }
        a.upd->norm_bits = mx;
    }
}

// ---- step 2: histogram limbs ------------------------------------------------------------------------------
__global__ void __launch_bounds__(FAB_THREADS)
k_fabric_hist(FabricArgs a) {
    const int par = a.epoch & 1;
    const int t = threadIdx.x;
    for (int r = 0; r < a.n; ++r) {
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.peer[r] + a.L.o_hist) +
                                  ((size_t)par * a.n + a.me) * HIST_WORDS;
        for (int i = t; i < HIST_WORDS; i += FAB_THREADS) dst[i] = a.hist[i];
    }
    fabric_signal_and_wait(a, 2);
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.peer[a.me] + a.L.o_hist) + (size_t)par * a.n * HIST_WORDS;
    for (int i = t; i < HIST_WORDS; i += FAB_THREADS) {
        unsigned long long s = 0ull;
        for (int r = 0; r < a.n; ++r) s += ld_cg(src + (size_t)r * HIST_WORDS + i);
        a.hist[i] = s;
    }
}

// ---- step 3: every shard's packed mask is complete --------------------------------------------------------
__global__ void __launch_bounds__(FAB_THREADS)
k_fabric_mask_ready(FabricArgs a) {
    fabric_signal_and_wait(a, 3);
}

}  // namespace boss
