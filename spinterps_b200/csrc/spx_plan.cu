// Host-side planning of one time chunk, native instead of NumPy: availability groups
// (interp/grps.py:57-101) and the descriptor arrays of the downdated kriging solves
// (spx_downdate).  Pure index work on the host -- the Python front end spent more time
// here (1250 groups per chunk of config 2) than the GPU needs for the chunk itself --
// plus the one small kernel that builds the right-hand-side matrix of the downdate from
// the resident data.
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <immintrin.h>
#include <mutex>
#include <sched.h>
#include <thread>
#include <unistd.h>
#include <vector>

#include "spx_b200.h"
#include "spx_common.cuh"
#include "spx_host_pool.h"

namespace spx {

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// One data row -> availability words (bit j % 64 of word j / 64 set where finite) and
// "any value >= thr" (NaN compares false, interp/steps.py:760-765).  SSE2 is the x86-64
// baseline; the AVX2 variant is chosen at run time.
static void scan_row_sse2(const double* __restrict__ src, int n, double thr,
                          uint64_t* __restrict__ bits, int* flag_io) {
    const __m128d thr2 = _mm_set1_pd(thr);
    int flag = *flag_io;
    for (int j0 = 0, w = 0; j0 < n; j0 += 64, ++w) {
        const int j1 = (n - j0 < 64) ? n - j0 : 64;
        uint64_t word = 0;
        int b = 0;
        for (; b + 2 <= j1; b += 2) {
            const __m128d v = _mm_loadu_pd(src + j0 + b);
            word |= (uint64_t)_mm_movemask_pd(_mm_cmpord_pd(v, v)) << b;
            flag |= _mm_movemask_pd(_mm_cmpge_pd(v, thr2));
        }
        for (; b < j1; ++b) {
            const double v = src[j0 + b];
            word |= (uint64_t)(v == v) << b;
            flag |= (int)(v >= thr);
        }
        bits[w] = word;
    }
    *flag_io = flag;
}

// AVX2: 8 doubles per iteration; the "any value >= thr" flag is a running maximum
// (max_pd keeps the accumulator when the new operand is NaN); `copy` (optional) receives
// the row in the same pass, with non-temporal stores when it is 32-byte aligned (a pinned
// staging buffer is written once and read by the DMA engine, never by this core).
__attribute__((target("avx2"))) static void scan_row_avx2(const double* __restrict__ src, int n,
                                                          double thr, uint64_t* __restrict__ bits,
                                                          int* flag_io, double* __restrict__ copy) {
    __m256d vmax = _mm256_set1_pd(-__builtin_inf());
    int any_inf_neg = 0;   // values equal to -inf compare >= thr only if thr is -inf itself
    const bool nt = copy && ((reinterpret_cast<uintptr_t>(copy) & 31) == 0);
    for (int j0 = 0, w = 0; j0 < n; j0 += 64, ++w) {
        const int j1 = (n - j0 < 64) ? n - j0 : 64;
        uint64_t word = 0;
        int b = 0;
        for (; b + 8 <= j1; b += 8) {
            const __m256d v0 = _mm256_loadu_pd(src + j0 + b);
            const __m256d v1 = _mm256_loadu_pd(src + j0 + b + 4);
            const unsigned m0 = (unsigned)_mm256_movemask_pd(_mm256_cmp_pd(v0, v0, _CMP_ORD_Q));
            const unsigned m1 = (unsigned)_mm256_movemask_pd(_mm256_cmp_pd(v1, v1, _CMP_ORD_Q));
            word |= (uint64_t)(m0 | (m1 << 4)) << b;
            vmax = _mm256_max_pd(v0, vmax);
            vmax = _mm256_max_pd(v1, vmax);
            if (copy) {
                if (nt) {
                    _mm256_stream_pd(copy + j0 + b, v0);
                    _mm256_stream_pd(copy + j0 + b + 4, v1);
                } else {
                    _mm256_storeu_pd(copy + j0 + b, v0);
                    _mm256_storeu_pd(copy + j0 + b + 4, v1);
                }
            }
        }
        for (; b < j1; ++b) {
            const double v = src[j0 + b];
            word |= (uint64_t)(v == v) << b;
            any_inf_neg |= (int)(v >= thr);
            if (copy) copy[j0 + b] = v;
        }
        bits[w] = word;
    }
    double m[4];
    _mm256_storeu_pd(m, vmax);
    const double mx = (m[0] > m[1] ? m[0] : m[1]) > (m[2] > m[3] ? m[2] : m[3])
                          ? (m[0] > m[1] ? m[0] : m[1]) : (m[2] > m[3] ? m[2] : m[3]);
    int flag = *flag_io | any_inf_neg;
    if (thr == -__builtin_inf()) {
        // every value that is not NaN counts: any available station (the running maximum
        // starts at -inf and cannot tell "no value" from "a value of -inf")
        for (int w = 0; w < (n + 63) / 64; ++w) flag |= (int)(bits[w] != 0);
    } else {
        flag |= (int)(mx >= thr);
    }
    *flag_io = flag;
}

// HostPool (spx_host_pool.h): persistent helper threads, leaked on purpose -- the detached
// workers wait on its condition variable until the process ends (a static object would be
// destroyed under them at exit).
struct HostPool::Impl {
    std::mutex mu, call_mu;
    std::condition_variable cv, done_cv;
    const std::function<void(int, int)>* fn = nullptr;
    uint64_t gen = 0;
    int pending = 0;
    int n_parts = 1;
    int n_workers = 0;
    int default_threads = 4;
    pid_t owner_pid = 0;

    void loop(int part) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)>* f;
            int parts;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return gen != seen; });
                seen = gen;
                f = fn;
                parts = n_parts;
            }
            if (part < parts) (*f)(part, parts);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) done_cv.notify_one();
            }
        }
    }
};

HostPool& HostPool::get() {
    static HostPool* p = new HostPool();
    return *p;
}

HostPool::HostPool() : impl_(new Impl()) {
    cpu_set_t set;
    int hw = 0;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) hw = CPU_COUNT(&set);
    // default: 4 threads, fewer when several ranks share the host (torchrun exports
    // LOCAL_WORLD_SIZE): half of this rank's share of the cores
    int n = 4;
    if (const char* lw = getenv("LOCAL_WORLD_SIZE")) {
        const int w = atoi(lw);
        if (w > 1 && hw > 0) {
            int share = hw / w / 2;
            if (share < 1) share = 1;
            if (share < n) n = share;
        }
    }
    if (const char* e = getenv("SPX_HOST_THREADS")) n = atoi(e);
    if (n < 1) n = 1;
    impl_->default_threads = n;
    int cap = 16;
    if (hw > 0 && cap > hw) cap = hw;
    if (n == 1 && getenv("SPX_HOST_THREADS")) cap = 1;       // SPX_HOST_THREADS=1: no helpers
    impl_->n_workers = cap - 1;
    impl_->owner_pid = getpid();
    for (int i = 0; i < impl_->n_workers; ++i) {
        Impl* im = impl_;
        std::thread([im, i] { im->loop(i + 1); }).detach();
    }
}

int HostPool::max_threads() const { return impl_->n_workers + 1; }

void HostPool::run(const std::function<void(int, int)>& fn, int n_threads) {
    Impl& im = *impl_;
    int parts = n_threads > 0 ? n_threads : im.default_threads;
    if (parts > im.n_workers + 1) parts = im.n_workers + 1;
    // a forked child inherits the object but not the worker threads
    if (parts <= 1 || getpid() != im.owner_pid) {
        fn(0, 1);
        return;
    }
    std::lock_guard<std::mutex> call_lock(im.call_mu);       // one job at a time
    {
        std::lock_guard<std::mutex> lk(im.mu);
        im.fn = &fn;
        im.n_parts = parts;
        im.pending = im.n_workers;
        ++im.gen;
    }
    im.cv.notify_all();
    fn(0, parts);
    std::unique_lock<std::mutex> lk(im.mu);
    im.done_cv.wait(lk, [&] { return im.pending == 0; });
    im.fn = nullptr;
}

// Bt rows of the downdate: row i < n_data is the data of step src_step[i] with NaN -> 0;
// row i >= n_data is the availability mask of step src_step[i] (1.0 where finite).
// Columns >= n_stn (the border) are zero.
__global__ void __launch_bounds__(256) k_build_bt(const double* __restrict__ data, int n_stn,
                                                  int64_t data_ld,
                                                  const int32_t* __restrict__ src_step,
                                                  int64_t n_rows, int64_t n_data, int M,
                                                  double* __restrict__ bt) {
    const int64_t row = blockIdx.x;
    if (row >= n_rows) return;
    const double* __restrict__ src = data + (int64_t)src_step[row] * data_ld;
    double* __restrict__ dst = bt + row * (int64_t)M;
    const bool mask_row = row >= n_data;
    for (int c = threadIdx.x; c < M; c += blockDim.x) {
        double v = 0.0;
        if (c < n_stn) {
            const double z = src[c];
            const bool fin = (z == z);
            v = mask_row ? (fin ? 1.0 : 0.0) : (fin ? z : 0.0);
        }
        dst[c] = v;
    }
}

// Ascending lists of the available (finite) and of the missing (NaN) stations of one
// step per warp (ballot / popc compaction): the kept / removed index sets of a downdated
// system, straight from the resident data.
__global__ void __launch_bounds__(256) k_avail_lists(const double* __restrict__ data, int n_stn,
                                                     int64_t data_ld,
                                                     const int32_t* __restrict__ src_step,
                                                     int n_sys,
                                                     const int64_t* __restrict__ stn_off,
                                                     int32_t* __restrict__ stn_list,
                                                     const int64_t* __restrict__ miss_off,
                                                     int32_t* __restrict__ miss_list) {
    const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n_sys) return;
    const double* __restrict__ src = data + (int64_t)src_step[w] * data_ld;
    int32_t* __restrict__ oa = stn_list + stn_off[w];
    int32_t* __restrict__ om = miss_list + miss_off[w];
    int na = 0, nm = 0;
    for (int c0 = 0; c0 < n_stn; c0 += 32) {
        const int c = c0 + lane;
        const bool in = c < n_stn;
        const double z = in ? src[c] : 0.0;
        const bool fin = in && (z == z);
        const unsigned ba = __ballot_sync(0xffffffffu, fin);
        const unsigned bm = __ballot_sync(0xffffffffu, in && !fin);
        const unsigned below = (1u << lane) - 1u;
        if (fin) oa[na + __popc(ba & below)] = c;
        else if (in) om[nm + __popc(bm & below)] = c;
        na += __popc(ba);
        nm += __popc(bm);
    }
}

}  // namespace spx

using namespace spx;

extern "C" {

int spx_avail_groups_host(const double* data, int64_t n_steps, int32_t n_stn, int64_t ld,
                          double min_var_thr, int32_t* grp_of_step, int32_t* grp_first,
                          int32_t* grp_n, uint8_t* grp_mask, int32_t* n_avail,
                          uint8_t* step_flag, double* data_copy, uint64_t* grp_bits_out,
                          int32_t* n_grps_out) {
    if (!data || !grp_of_step || !grp_first || !grp_n || !n_avail || !step_flag ||
        !n_grps_out || n_steps < 0 || n_stn < 1 || ld < n_stn) {
        set_error("avail_groups_host: bad argument");
        return SPX_EINVAL;
    }
    const int W = (n_stn + 63) / 64;
    // ---- pass 1 (parallel over rows): availability words, count, flag, hash, copy -------
    static thread_local std::vector<uint64_t> row_bits;
    static thread_local std::vector<uint64_t> row_hash;
    row_bits.resize((size_t)n_steps * W);
    row_hash.resize((size_t)n_steps);
    uint64_t* rb = row_bits.data();
    uint64_t* rh = row_hash.data();
    const bool use_avx2 = __builtin_cpu_supports("avx2") != 0;
    auto scan = [&](int part, int n_parts) {
        const int64_t t0 = n_steps * part / n_parts, t1 = n_steps * (part + 1) / n_parts;
        for (int64_t t = t0; t < t1; ++t) {
            const double* __restrict__ src = data + t * ld;
            uint64_t* __restrict__ bits = rb + t * W;
            int flag = 0;
            if (use_avx2) {
                scan_row_avx2(src, n_stn, min_var_thr, bits, &flag,
                              data_copy ? data_copy + t * (int64_t)n_stn : nullptr);
            } else {
                scan_row_sse2(src, n_stn, min_var_thr, bits, &flag);
                if (data_copy)
                    std::memcpy(data_copy + t * (int64_t)n_stn, src, sizeof(double) * n_stn);
            }
            int cnt = 0;
            uint64_t h = 0x9e3779b97f4a7c15ULL;
            for (int w = 0; w < W; ++w) {
                cnt += __builtin_popcountll(bits[w]);
                h = mix64(h ^ bits[w]);
            }
            rh[t] = h;
            n_avail[t] = cnt;
            step_flag[t] = (uint8_t)(flag != 0);
        }
        if (data_copy && use_avx2) _mm_sfence();       // non-temporal stores visible to the DMA
    };
    if (n_steps * (int64_t)n_stn >= (1 << 16))
        HostPool::get().run(scan);
    else
        scan(0, 1);
    // ---- pass 2 (sequential): groups in first-occurrence order ---------------------------
    std::vector<uint64_t> grp_bits;                    // [n_grps, W]
    grp_bits.reserve((size_t)W * 64);
    size_t cap = 16;
    while (cap < (size_t)n_steps * 2 + 2) cap <<= 1;
    std::vector<int32_t> table(cap, -1);
    int32_t n_grps = 0;
    for (int64_t t = 0; t < n_steps; ++t) {
        const uint64_t* bits = rb + t * W;
        const int cnt = n_avail[t];
        size_t slot = (size_t)rh[t] & (cap - 1);
        int32_t g = -1;
        for (;;) {
            const int32_t cand = table[slot];
            if (cand < 0) break;
            if (std::memcmp(&grp_bits[(size_t)cand * W], bits, sizeof(uint64_t) * W) == 0) {
                g = cand;
                break;
            }
            slot = (slot + 1) & (cap - 1);
        }
        if (g < 0) {
            g = n_grps++;
            table[slot] = g;
            grp_bits.insert(grp_bits.end(), bits, bits + W);
            grp_first[g] = (int32_t)t;
            grp_n[g] = cnt;
            if (grp_bits_out)
                std::memcpy(grp_bits_out + (int64_t)g * W, bits, sizeof(uint64_t) * W);
            if (grp_mask) {
                uint8_t* __restrict__ m = grp_mask + (int64_t)g * n_stn;
                int j = 0;
                for (; j + 8 <= n_stn; j += 8) {
                    // 8 bits -> 8 bytes of 0 / 1 (byte i <- bit i)
                    const uint64_t b = (bits[j >> 6] >> (j & 63)) & 0xffu;
                    const uint64_t x = (b * 0x0101010101010101ULL) & 0x8040201008040201ULL;
                    const uint64_t y = ((x + 0x7f7f7f7f7f7f7f7fULL) >> 7) & 0x0101010101010101ULL;
                    std::memcpy(m + j, &y, 8);
                }
                for (; j < n_stn; ++j) m[j] = (uint8_t)((bits[j >> 6] >> (j & 63)) & 1u);
            }
        }
        grp_of_step[t] = g;
    }
    *n_grps_out = n_grps;
    return SPX_OK;
}

static inline int64_t align16(int64_t x) { return (x + 15) & ~(int64_t)15; }

int64_t spx_downdate_plan_bytes(int64_t n_sel, int32_t n_stn) {
    if (n_sel < 0 || n_stn < 0) return 0;
    // per system: both station lists (r + n = n_stn ints, filled on the device) and 9
    // descriptor entries; per right-hand side (<= 2 per selected step): urow, row, kind,
    // bt_step
    return 64 * 16 + n_sel * ((int64_t)n_stn * 4 + 96) + 2 * n_sel * 24;
}

int64_t spx_downdate_plan_host_bytes(int64_t n_sel) {
    if (n_sel < 0) return 0;
    return 64 * 16 + n_sel * 96 + 2 * n_sel * 24;     // the host-written prefix only
}

int spx_downdate_plan_host(const int32_t* grp_of_step, const int32_t* grp_n, int32_t n_grps,
                           int32_t n_stn, const int32_t* steps, const int64_t* rows,
                           int64_t n_sel, void* buf, int64_t buf_bytes, spx_dd_plan* plan) {
    if (!grp_of_step || !grp_n || !steps || !rows || !buf || !plan || n_grps < 1 || n_stn < 1 ||
        n_sel < 1) {
        set_error("downdate_plan_host: bad argument");
        return SPX_EINVAL;
    }
    if (buf_bytes < spx_downdate_plan_host_bytes(n_sel)) {
        set_error("downdate_plan_host: buffer of %lld bytes, need %lld", (long long)buf_bytes,
                  (long long)spx_downdate_plan_host_bytes(n_sel));
        return SPX_ENOMEM;
    }
    // systems = groups that occur among the selected steps, ascending group id
    std::vector<int32_t> cnt((size_t)n_grps, 0);
    for (int64_t i = 0; i < n_sel; ++i) {
        const int32_t g = grp_of_step[steps[i]];
        if (g < 0 || g >= n_grps) {
            set_error("downdate_plan_host: group id out of range");
            return SPX_EINVAL;
        }
        ++cnt[g];
    }
    std::vector<int32_t> sys_of_grp((size_t)n_grps, -1);
    int32_t n_sys = 0;
    int64_t total_r = 0, total_n = 0;
    int32_t max_r = 0;
    for (int32_t g = 0; g < n_grps; ++g)
        if (cnt[g]) {
            if (grp_n[g] < 0 || grp_n[g] > n_stn) {
                set_error("downdate_plan_host: grp_n out of range");
                return SPX_EINVAL;
            }
            sys_of_grp[g] = n_sys++;
            const int32_t r = n_stn - grp_n[g];
            total_r += r;
            total_n += grp_n[g];
            if (r > max_r) max_r = r;
        }
    const int64_t n_data = n_sel, n_rhs = n_sel + n_sys;

    uint8_t* base = static_cast<uint8_t*>(buf);
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        const int64_t o = off;
        off = align16(off + bytes);
        return o;
    };
    std::memset(plan, 0, sizeof(*plan));
    plan->n_sys = n_sys;
    plan->n_data = (int32_t)n_data;
    plan->n_rhs = (int32_t)n_rhs;
    plan->max_r = max_r;
    plan->total_r = total_r;
    plan->total_n = total_n;
    plan->off_sys_r = take(4 * (int64_t)n_sys);
    plan->off_sys_miss_off = take(8 * (int64_t)n_sys);
    plan->off_sys_n = take(4 * (int64_t)n_sys);
    plan->off_sys_stn_off = take(8 * (int64_t)n_sys);
    plan->off_sys_rhs_off = take(8 * (int64_t)n_sys);
    plan->off_sys_rhs_cnt = take(4 * (int64_t)n_sys);
    plan->off_rhs_urow = take(4 * n_rhs);
    plan->off_rhs_row = take(8 * n_rhs);
    plan->off_rhs_kind = take(4 * n_rhs);
    plan->off_sys_order = take(4 * (int64_t)n_sys);
    plan->off_bt_step = take(4 * n_rhs);
    plan->off_sys_grp = take(4 * (int64_t)n_sys);
    plan->off_pos_ones = take(8 * (int64_t)n_sys);
    plan->n_upload_bytes = off;
    // device-filled tail (spx_avail_lists_dev): never written on the host
    plan->off_miss_list = take(4 * (total_r > 0 ? total_r : 1));
    plan->off_stn_list = take(4 * (total_n > 0 ? total_n : 1));
    plan->n_bytes = off;
    if (plan->n_upload_bytes > buf_bytes) {
        set_error("downdate_plan_host: plan needs %lld bytes", (long long)plan->n_upload_bytes);
        return SPX_ENOMEM;
    }
    auto* sys_r = reinterpret_cast<int32_t*>(base + plan->off_sys_r);
    auto* sys_miss_off = reinterpret_cast<int64_t*>(base + plan->off_sys_miss_off);
    auto* sys_n = reinterpret_cast<int32_t*>(base + plan->off_sys_n);
    auto* sys_stn_off = reinterpret_cast<int64_t*>(base + plan->off_sys_stn_off);
    auto* sys_rhs_off = reinterpret_cast<int64_t*>(base + plan->off_sys_rhs_off);
    auto* sys_rhs_cnt = reinterpret_cast<int32_t*>(base + plan->off_sys_rhs_cnt);
    auto* rhs_urow = reinterpret_cast<int32_t*>(base + plan->off_rhs_urow);
    auto* rhs_row = reinterpret_cast<int64_t*>(base + plan->off_rhs_row);
    auto* rhs_kind = reinterpret_cast<int32_t*>(base + plan->off_rhs_kind);
    auto* sys_order = reinterpret_cast<int32_t*>(base + plan->off_sys_order);
    auto* bt_step = reinterpret_cast<int32_t*>(base + plan->off_bt_step);
    auto* sys_grp = reinterpret_cast<int32_t*>(base + plan->off_sys_grp);
    auto* pos_ones = reinterpret_cast<int64_t*>(base + plan->off_pos_ones);

    // right-hand-side segments: the data rows of a system in the order of `steps`, then
    // its ones-vector
    int64_t mo = 0, so = 0, ro = 0, uo = 0;
    std::vector<int64_t> fill((size_t)n_sys);     // next rhs slot of each system
    std::vector<int64_t> ufill((size_t)n_sys);    // next Bt data row of each system
    for (int32_t g = 0; g < n_grps; ++g) {
        const int32_t s = sys_of_grp[g];
        if (s < 0) continue;
        sys_grp[s] = g;
        const int32_t n = grp_n[g], r = n_stn - n;
        sys_r[s] = r;
        sys_n[s] = n;
        sys_miss_off[s] = mo;
        sys_stn_off[s] = so;
        mo += r;
        so += n;
        sys_rhs_off[s] = ro;
        sys_rhs_cnt[s] = cnt[g] + 1;
        fill[s] = ro;
        ufill[s] = uo;
        pos_ones[s] = ro + cnt[g];
        rhs_urow[ro + cnt[g]] = (int32_t)(n_data + s);
        rhs_row[ro + cnt[g]] = -1;
        rhs_kind[ro + cnt[g]] = 1;
        ro += cnt[g] + 1;
        uo += cnt[g];
    }
    for (int64_t i = 0; i < n_sel; ++i) {
        const int32_t s = sys_of_grp[grp_of_step[steps[i]]];
        const int64_t q = fill[s]++;
        const int64_t u = ufill[s]++;
        rhs_urow[q] = (int32_t)u;
        rhs_row[q] = rows[i];
        rhs_kind[q] = 0;
        bt_step[u] = steps[i];
        // the mask row of the system comes from any of its steps (all share the mask)
        bt_step[n_data + s] = steps[i];
    }
    // largest systems first (stable counting sort on r, descending) shortens the tail
    {
        std::vector<int32_t> start((size_t)max_r + 2, 0);
        for (int32_t s = 0; s < n_sys; ++s) ++start[(size_t)(max_r - sys_r[s]) + 1];
        for (int32_t k = 0; k <= max_r; ++k) start[(size_t)k + 1] += start[(size_t)k];
        for (int32_t s = 0; s < n_sys; ++s) sys_order[start[(size_t)(max_r - sys_r[s])]++] = s;
    }
    return SPX_OK;
}

int spx_avail_lists_dev(const double* data, int32_t n_stn, int64_t data_ld,
                        const int32_t* src_step, int32_t n_sys, const int64_t* stn_off,
                        int32_t* stn_list, const int64_t* miss_off, int32_t* miss_list,
                        void* stream) {
    if (n_sys == 0) return SPX_OK;
    if (!data || !src_step || !stn_off || !stn_list || !miss_off || !miss_list || n_stn < 1) {
        set_error("avail_lists: bad argument");
        return SPX_EINVAL;
    }
    const int64_t threads = (int64_t)n_sys * 32;
    k_avail_lists<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        data, n_stn, data_ld, src_step, n_sys, stn_off, stn_list, miss_off, miss_list);
    SPX_CHECK_LAUNCH("k_avail_lists");
    return SPX_OK;
}

int spx_build_bt_dev(const double* data, int32_t n_stn, int64_t data_ld, const int32_t* src_step,
                     int64_t n_rows, int64_t n_data, int32_t n_border, double* bt, void* stream) {
    if (n_rows == 0) return SPX_OK;
    if (!data || !src_step || !bt || n_stn < 1 || n_border < 0 || n_data > n_rows) {
        set_error("build_bt: bad argument");
        return SPX_EINVAL;
    }
    k_build_bt<<<(unsigned)n_rows, 256, 0, (cudaStream_t)stream>>>(
        data, n_stn, data_ld, src_step, n_rows, n_data, n_stn + n_border, bt);
    SPX_CHECK_LAUNCH("k_build_bt");
    return SPX_OK;
}

}  // extern "C"
