// C-ABI of libslicq (include/slicq.h): plan construction and the chunked launch sequences.
//
// forward  : per chunk of (row,slice) units   slice_fft_fwd_kernel -> bins_fwd_kernel
// inverse  : per chunk                         bins_inv_kernel -> slice_fft_inv_kernel (even slices, then odd slices)
// A chunk's intermediate spectra (analysis: padded half spectra H; synthesis: the two bin planes T) live in the
// caller-provided scratch buffer; SLICQ_CHUNK_MB bounds the chunk (default: one chunk per call).
#include "slicq_common.cuh"
#include "../../include/slicq.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <atomic>
#include <mutex>

#ifdef SLICQ_EMU
#include <condition_variable>
#include <mutex>
#include <thread>
thread_local dim3 threadIdx;
dim3 blockIdx, blockDim, gridDim;
static unsigned char slicq_emu_smem_buf[256 * 1024] __attribute__((aligned(16)));
unsigned char* slicq_emu_smem = slicq_emu_smem_buf;
namespace {
struct EmuBarrier {
    std::mutex m; std::condition_variable cv; int count = 0, gen = 0, n = 1;
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const int g = gen;
        if (++count == n) { count = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
} g_emu_bar;
}  // namespace
void slicq_emu_sync() { g_emu_bar.wait(); }
void slicq_emu_launch(dim3 grid, dim3 block, const std::function<void()>& body) {
    gridDim = grid; blockDim = block; g_emu_bar.n = (int)block.x;
    std::vector<std::thread> th(block.x);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = dim3(bx, by, bz);
                for (unsigned t = 0; t < block.x; ++t)
                    th[t] = std::thread([t, &body]() { threadIdx = dim3(t, 0, 0); body(); });
                for (unsigned t = 0; t < block.x; ++t) th[t].join();
            }
}
#endif

extern "C" int slicq_launch_bins(const SlicqBinsParams* p, int n_tiles, int smem_bytes, int synth, cudaStream_t s);
extern "C" int slicq_launch_slice_fwd(const SlicqSliceParams* p, cudaStream_t s);
extern "C" int slicq_launch_slice_inv(const SlicqSliceParams* p, cudaStream_t s);
extern "C" int slicq_slice_smem_bytes(int L);
extern "C" int slicq_bins_threads(void);
extern "C" int slicq_launch_mirror_fix(const SlicqBinsParams* p, cudaStream_t s);

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

// ---- optional per-kernel CUDA-event timing (bench.py roofline accounting) -----------------
enum { K_SLICE_FWD = 0, K_BINS_FWD, K_BINS_INV, K_SLICE_INV, K_COUNT };
std::atomic<bool> g_prof{false};
#ifndef SLICQ_EMU
struct ProfRec { int kid; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof_recs;
std::mutex g_prof_mutex;
#endif
struct ProfScope {
#ifndef SLICQ_EMU
    int kid; cudaStream_t s; cudaEvent_t a, b; bool on;
    ProfScope(int k, cudaStream_t st) : kid(k), s(st), on(g_prof) {
        if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); }
    }
    ~ProfScope() {
        if (on) { cudaEventRecord(b, s); ProfRec r; r.kid = kid; r.a = a; r.b = b; std::lock_guard<std::mutex> lk(g_prof_mutex); g_prof_recs.push_back(r); }
    }
#else
    ProfScope(int, cudaStream_t) {}
#endif
};

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

struct FftPlan { int M, kind, A, B, cost; };
const FftPlan kFftPlans[] = {
#define SLICQ_FFT_SIZE(M_, K_, A_, B_, C_) {M_, K_, A_, B_, C_},
#include "fft_sizes.inc"
#undef SLICQ_FFT_SIZE
};

const FftPlan* find_fft_plan(int M) {
    for (size_t i = 0; i < sizeof(kFftPlans) / sizeof(kFftPlans[0]); ++i)
        if (kFftPlans[i].M == M) return &kFftPlans[i];
    return nullptr;
}

// shared-memory bytes one transform of this plan needs (max over analysis / synthesis)
int fft_smem_per_transform(const FftPlan& f) {
    if (f.kind == 1) return (f.M + 1) * 8;
    if (f.kind == 2) {
        const int bp = (f.B % 2 == 0) ? f.B + 1 : f.B;
        const int ap = (f.A % 2 == 0) ? f.A + 1 : f.A;
        const int a = f.A * bp, b = f.B * ap;
        return (a > b ? a : b) * 8;
    }
    return (f.M + f.A) * 8;  // kind 3: [R][P] (synthesis) / [P][R + 1] (analysis) complex
}

// threads one transform occupies in the pass whose per-thread state is hoisted out of the unit loop
int fft_threads_per_transform(const FftPlan& f) {
    if (f.kind == 1) return 1;
    if (f.kind == 2) return f.B;
    return f.B * (f.A <= 41 ? 2 : (f.A <= 61 ? 3 : 4));  // kind 3: (part, n2): the DFT-P is split by outputs over 2-4 threads
}

struct Bucket {
    int M, first_bin, n_bins, gt, tw_off, kind, A, B, smem_per_fft;
    int gap_first, gap_n;   // entries of the gap list (zero-filled positions of T) this bucket owns
    double cost;     // relative work per unit (for the job split)
};

template <class T> int upload(const std::vector<T>& h, const T** d, std::vector<void*>& owned) {
    void* p = nullptr;
    if (cudaMalloc(&p, h.size() * sizeof(T)) != cudaSuccess) return -1;
    if (cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    owned.push_back(p);
    *d = reinterpret_cast<const T*>(p);
    return 0;
}

}  // namespace

#define SLICQ_MAX_WAYS 8
struct slicq_plan {
    int L, N2, hop, J, sum_M;
    std::vector<int> bin_M, bin_pos, bin_coff;
    std::vector<Bucket> buckets;
    SlicqDeviceTables dev;
    std::vector<void*> owned;
    int bins_smem;        // dynamic shared memory of the bins kernels
    int pad_l, pad_r;
    long long spec_stride_fwd;
    long long t_stride;   // complex elements per unit of the synthesis intermediate T
    long long chunk_bytes;
    int target_jobs;      // CTAs per bins launch the job split aims for
    int min_iters;        // ... but a job keeps at least this many iterations (instruction-cache reuse)
    int only_bucket;      // -1, or (SLICQ_ONLY_BUCKET, tuning aid) the single bucket the bins kernels process
    // Large calls are split by rows into two halves that run on two internal streams (forked from and
    // joined back into the caller's stream): the memory-latency-bound and the issue-bound kernels of
    // the two halves overlap on the SMs.  Rows are independent, so the results do not change.
    long long split_units;          // split when the call has at least this many units (0 = never)
    int split_ways;                 // row groups / internal streams of a split call (2 .. SLICQ_MAX_WAYS)
    mutable cudaStream_t side[SLICQ_MAX_WAYS];
    mutable cudaEvent_t ev_fork, ev_join[SLICQ_MAX_WAYS];
    mutable bool side_ready;
    // calls that split share the side streams and the fork / join events: their enqueue sequence (record, waits, launches,
    // joins) runs under this mutex, so that one thread's waits always see its own records.  Un-split calls take no lock.
    mutable std::mutex split_mutex;
};

extern "C" int slicq_abi_version(void) { return SLICQ_ABI_VERSION; }
#ifdef SLICQ_EMU
extern "C" int slicq_build_kind(void) { return 1; }
#else
extern "C" int slicq_build_kind(void) { return 0; }
#endif
extern "C" const char* slicq_last_error(void) { return g_err.c_str(); }
extern "C" int64_t slicq_launch_count(void) { return g_launches.load(); }

extern "C" int slicq_profile_enable(int on) { g_prof = on != 0; return SLICQ_OK; }

extern "C" int slicq_profile_read(double* ms, int64_t* launches) {
    for (int i = 0; i < K_COUNT; ++i) { if (ms) ms[i] = 0.0; if (launches) launches[i] = 0; }
#ifndef SLICQ_EMU
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    for (ProfRec& r : g_prof_recs) {
        float t = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            if (ms) ms[r.kid] += t;
            if (launches) launches[r.kid] += 1;
        }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof_recs.clear();
#endif
    return SLICQ_OK;
}

extern "C" void slicq_plan_destroy(slicq_plan* p) {
    if (!p) return;
    if (p->side_ready) {
        for (int h = 0; h < SLICQ_MAX_WAYS; ++h) { cudaStreamDestroy(p->side[h]); cudaEventDestroy(p->ev_join[h]); }
        cudaEventDestroy(p->ev_fork);
    }
    for (void* q : p->owned) cudaFree(q);
    delete p;
}

extern "C" int slicq_plan_create(const slicq_tables* t, slicq_plan** out) {
    if (!t || !out) return fail(SLICQ_E_INVALID, "null argument");
    *out = nullptr;
    const int L = t->sl_len, J = t->n_bins;
    if (L <= 0 || L % 4 != 0) return fail(SLICQ_E_INVALID, "sl_len must be a positive multiple of 4");
    if (J < 2 || !t->bin_M || !t->bin_pos || !t->win_fwd || !t->win_inv || !t->tukey)
        return fail(SLICQ_E_INVALID, "missing tables");
    if (slicq_slice_smem_bytes(L) < 0) {
        char buf[160];
        snprintf(buf, sizeof buf, "slice length %d does not fit the slice kernels' shared memory (two buffers of sl_len/2 complex numbers)", L);
        return fail(SLICQ_E_UNSUPPORTED, buf);
    }
    slicq_plan* p = new slicq_plan();
    p->L = L; p->N2 = L / 2; p->hop = L / 2; p->J = J;
    p->bin_M.assign(t->bin_M, t->bin_M + J);
    p->bin_pos.assign(t->bin_pos, t->bin_pos + J);
    p->bin_coff.resize(J);
    int off = 0;
    for (int j = 0; j < J; ++j) {
        const int M = p->bin_M[j], pos = p->bin_pos[j];
        if (M < 4 || M % 4 != 0 || M > SLICQ_MAX_M || !find_fft_plan(M)) {
            delete p;
            return fail(SLICQ_E_UNSUPPORTED, "bin length M must be a multiple of 4 in [16, 292]");
        }
        if (pos % 2 != 0 || pos < 0 || pos > p->N2 || (j > 0 && pos <= p->bin_pos[j - 1])) {
            delete p;
            return fail(SLICQ_E_UNSUPPORTED, "bin positions must be even, increasing and within [0, L/2]");
        }
        p->bin_coff[j] = off;
        off += M;
    }
    p->sum_M = off;
    // The reference also accumulates a mirrored copy of bins 1..J-2 at L - pos_j (nsigtf.py:63-80).  It reaches the kept
    // half spectrum [0, L/2] only from a bin that extends past Nyquist (unsupported) or below DC (handled by
    // mirror_fix_kernel, entries built below).
    for (int j = 1; j < J - 1; ++j) {
        if (p->bin_pos[j] + p->bin_M[j] / 2 > p->N2) {
            delete p;
            return fail(SLICQ_E_UNSUPPORTED, "a bin below Nyquist reaches beyond it: the mirrored-bin pass above Nyquist is not implemented");
        }
    }
    // buckets = maximal runs of equal M (nsgtf.py:66-78)
    int tw_off = 0;
    for (int j = 0; j < J; ++j) {
        if (!p->buckets.empty() && p->buckets.back().M == p->bin_M[j]) {
            p->buckets.back().n_bins++;
            continue;
        }
        const FftPlan* f = find_fft_plan(p->bin_M[j]);
        Bucket b;
        b.M = p->bin_M[j]; b.first_bin = j; b.n_bins = 1; b.tw_off = tw_off;
        b.kind = f->kind; b.A = f->A; b.B = f->B; b.smem_per_fft = fft_smem_per_transform(*f);
        b.gt = 1; b.cost = 0.0; b.gap_first = 0; b.gap_n = 0;
        tw_off += b.M;
        p->buckets.push_back(b);
    }
    if ((int)p->buckets.size() > SLICQ_MAX_BUCKETS) {
        delete p;
        return fail(SLICQ_E_UNSUPPORTED, "too many buckets");
    }
    // units per CTA iteration: fill the 256 threads of the hoisted pass, shared memory <= 48 KB
    p->bins_smem = 0;
    for (Bucket& b : p->buckets) {
        const FftPlan* f = find_fft_plan(b.M);
        const int nthr = slicq_bins_threads();
        int gt = nthr / (b.n_bins * fft_threads_per_transform(*f));
        if (gt < 1) gt = 1;
        static const int stage_kb = getenv("SLICQ_BINS_STAGE_KB") ? atoi(getenv("SLICQ_BINS_STAGE_KB")) : 48;   // tuning aid
        while (gt > 1 && b.n_bins * gt * b.smem_per_fft > stage_kb * 1024) --gt;
        if (b.n_bins * fft_threads_per_transform(*f) > nthr) {
            delete p;
            return fail(SLICQ_E_UNSUPPORTED, "bucket has too many bins for one CTA");
        }
        b.gt = gt;
        // + SLICQ_SLOT_BYTES; single-thread transforms keep the bucket's windows behind the stage
        // two-pass and prime transforms keep the job's twiddles (M complex) and dual windows (F * M floats) there too
        const int sm = b.n_bins * gt * b.smem_per_fft + 4096 + (f->kind == 1 ? b.n_bins * b.M * 4 + b.n_bins * 12 + 32 : 0) + (f->kind == 3 ? b.n_bins * 12 + 16 : 0) +
                       (f->kind >= 2 ? b.M * 8 + b.n_bins * b.M * 4 + b.n_bins * 4 + 16 : 0);
        if (sm > p->bins_smem) p->bins_smem = sm;
        // work model: flops ~ M log2 M per transform plus a per-coefficient load/store term
        b.cost = (double)b.n_bins * f->cost;   // instructions per transform (fft_sizes.inc)
    }

    // ---- derived tables
    std::vector<float> wf(p->sum_M), wi(p->sum_M), tuk(t->tukey, t->tukey + L);
    int pad_l = 0, pad_r = 0;
    for (int j = 0; j < J; ++j) {
        const int M = p->bin_M[j], o = p->bin_coff[j];
        const double sgn = ((p->bin_pos[j] / 2) % 2) ? -1.0 : 1.0;
        for (int mc = 0; mc < M; ++mc) {            // centred order m' = m~ + M/2
            const int m = (mc + M / 2) % M;         // reference order (peak at m = 0)
            wf[o + mc] = (float)((double)t->win_fwd[o + m] * sgn / (double)M);
            wi[o + mc] = (float)((double)t->win_inv[o + m] * sgn * (double)M / (double)L);   // 1/L of the slice IFFT folded in
        }
        pad_l = std::max(pad_l, M / 2 - p->bin_pos[j]);
        pad_r = std::max(pad_r, p->bin_pos[j] + M / 2 - 1 - p->N2);
    }
    pad_l = (pad_l + 1) & ~1;
    pad_r = (pad_r + 1) & ~1;
    if (pad_l > p->N2 / 2 || pad_r > p->N2 / 2) {
        delete p;
        return fail(SLICQ_E_UNSUPPORTED, "bins reach too far beyond DC / Nyquist");
    }
    p->pad_l = pad_l; p->pad_r = pad_r;
    int tw_lo = 0, tw_hi = L;
    while (tw_lo < L && tuk[tw_lo] == 0.f) ++tw_lo;
    while (tw_hi > tw_lo && tuk[tw_hi - 1] == 0.f) --tw_hi;
    std::vector<float2> post(p->N2 / 2 + 1), tw(tw_off);
    for (int k = 0; k <= p->N2 / 2; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)L;
        post[k] = make_float2((float)cos(a), (float)sin(a));
    }
    for (const Bucket& b : p->buckets)
        for (int j = 0; j < b.M; ++j) {
            const double a = -2.0 * M_PI * (double)j / (double)b.M;
            tw[b.tw_off + j] = make_float2((float)cos(a), (float)sin(a));
        }
    // synthesis intermediate T (see SlicqDeviceTables): two planes + overflow + gap list
    const int N2 = p->N2;
    const int pl_off = pad_l;                                            // even
    const int pl_len = (pl_off + N2 + 2 + std::max(pad_r, 2) + 1) & ~1;  // positions -pl_off .. N2 + 1 + pad_r, even
    std::vector<int> toff(J), ov(J, 0), ovoff(J, 0);
    int n_ovf = 0;
    for (int j = 0; j < J; ++j) {
        const int lo = p->bin_pos[j] - p->bin_M[j] / 2;
        toff[j] = (j & 1) * pl_len + pl_off + lo;
        if (j >= 2) {
            const int end_prev = p->bin_pos[j - 2] + p->bin_M[j - 2] / 2;
            ov[j] = std::max(0, end_prev - lo);
        }
        if (ov[j] > p->bin_M[j] || (ov[j] & 1)) { delete p; return fail(SLICQ_E_UNSUPPORTED, "unsupported overlap between bins of one plane"); }
        if (j >= 4 && p->bin_pos[j - 4] + p->bin_M[j - 4] / 2 > lo) { delete p; return fail(SLICQ_E_UNSUPPORTED, "more than 4 bins overlap at one spectrum position"); }
        ovoff[j] = 2 * pl_len + n_ovf;
        n_ovf += ov[j];
    }
    // two-pass / prime transforms store their first run of A outputs through one select: the overflow part must fit
    for (const Bucket& b : p->buckets)
        for (int j = b.first_bin; j < b.first_bin + b.n_bins; ++j)
            if ((b.kind == 2 && ov[j] > b.A) || (b.kind == 3 && ov[j] > b.B)) { delete p; return fail(SLICQ_E_UNSUPPORTED, "overlap between bins of one plane exceeds the first transform pass"); }
    p->t_stride = (2LL * pl_len + n_ovf + 1) & ~1LL;
    std::vector<int4> ex;
    {
        std::vector<int> e0(N2 + 1, -1), e1(N2 + 1, -1);
        for (int j = 0; j < J; ++j)
            for (int m = 0; m < ov[j]; ++m) {
                const int f = p->bin_pos[j] - p->bin_M[j] / 2 + m;
                if (f < 0 || f > N2) continue;
                if (e0[f] < 0) e0[f] = ovoff[j] + m; else e1[f] = ovoff[j] + m;
            }
        for (int f = 0; f <= N2; ++f)
            if (e0[f] >= 0) ex.push_back(int4{f, e0[f], e1[f], 0});
    }
    const int n_ex = (int)ex.size();
    if (ex.empty()) ex.push_back(int4{0, 0, -1, 0});
    // gaps: positions of a plane row (-pl_off .. pl_len - pl_off - 1, margins included: the adjoint-of-analysis mode folds
    // them back) that none of its bins covers; owner = the next bin of the plane (or its last)
    std::vector<int2> gaps;
    {
        std::vector<std::vector<int2> > per_bucket(p->buckets.size());
        auto bucket_of = [&](int j) { for (size_t i = 0; i < p->buckets.size(); ++i) if (j >= p->buckets[i].first_bin && j < p->buckets[i].first_bin + p->buckets[i].n_bins) return (int)i; return 0; };
        for (int q = 0; q < 2; ++q) {
            std::vector<int> owner(pl_len, -1);    // covering bin or -1, index = pl_off + f
            int last = -1;
            for (int j = q; j < J; j += 2) {
                const int lo = std::max(0, pl_off + p->bin_pos[j] - p->bin_M[j] / 2), hi = std::min(pl_len, pl_off + p->bin_pos[j] + p->bin_M[j] / 2);
                for (int f = lo; f < hi; ++f) owner[f] = j;
                last = j;
            }
            if (last < 0) { delete p; return fail(SLICQ_E_UNSUPPORTED, "a bin plane is empty"); }
            for (int f = 0; f < pl_len;) {
                if (owner[f] >= 0) { ++f; continue; }
                int g = f;
                while (g < pl_len && owner[g] < 0) ++g;
                const int next = g < pl_len ? owner[g] : last;
                per_bucket[bucket_of(next)].push_back(int2{q * pl_len + f, g - f});
                f = g;
            }
        }
        for (size_t i = 0; i < p->buckets.size(); ++i) {
            p->buckets[i].gap_first = (int)gaps.size();
            p->buckets[i].gap_n = (int)per_bucket[i].size();
            gaps.insert(gaps.end(), per_bucket[i].begin(), per_bucket[i].end());
        }
    }
    if (gaps.empty()) gaps.push_back(int2{0, 0});
    // mirrored-bin entries (see mirror_fix_kernel): bin j >= 1 with pos_j < M_j / 2; reference-order index m in [pos_j, M_j/2)
    std::vector<SlicqMirrorEntry> mir;
    for (size_t bi = 0; bi < p->buckets.size(); ++bi) {
        const Bucket& b = p->buckets[bi];
        for (int j = std::max(1, b.first_bin); j < std::min(J - 1, b.first_bin + b.n_bins); ++j) {
            const int M = b.M, pos = p->bin_pos[j];
            for (int m = pos; m < M / 2; ++m) {
                const int f = m - pos;
                if (f > N2) continue;
                SlicqMirrorEntry e;
                e.bucket = (int)bi; e.f_in_bucket = j - b.first_bin; e.m_src = m + 1; e.t_off = pl_off + f;
                const double gdm = (double)t->win_inv[p->bin_coff[j] + (M - m) % M];      // dual window of the mirrored bin at m
                const double sgn = (((pos / 2) + m) % 2) ? -1.0 : 1.0;                     // (-1)^(pos/2) (-1)^m
                e.weight = (float)(sgn * gdm * (double)M / (double)L);
                mir.push_back(e);
            }
        }
    }
    // (the adjoint-of-analysis plan has no such pass: the analysis reads the exact Hermitian mirror)
    const int n_mir = (getenv("SLICQ_NO_MIRROR") || (t->flags & SLICQ_PLAN_ADJOINT_OF_ANALYSIS)) ? 0 : (int)mir.size();   // SLICQ_NO_MIRROR: verification aid (shows what the pass contributes)
    if (mir.empty()) mir.push_back(SlicqMirrorEntry{0, 0, 0, 0, 0.f});
    // generic slice kernels: prime factors of N2 (largest first) and the table exp(-2 pi i k / N2)
    std::vector<int> fac;
    { int n = N2; for (int d = 2; d * d <= n; ++d) while (n % d == 0) { fac.push_back(d); n /= d; } if (n > 1) fac.push_back(n); }
    std::sort(fac.begin(), fac.end(), [](int x, int y) { return x > y; });
    if (fac.size() > 20) { delete p; return fail(SLICQ_E_UNSUPPORTED, "too many prime factors in the slice length"); }
    std::vector<double2> wP(fac[0] > 32 ? fac[0] : 1);
    for (size_t kk = 0; kk < wP.size(); ++kk) {
        const double ang = -2.0 * M_PI * (double)kk / (double)wP.size();
        wP[kk].x = cos(ang); wP[kk].y = sin(ang);
    }
    std::vector<float2> wN(N2);
    for (int kk = 0; kk < N2; ++kk) {
        const double ang = -2.0 * M_PI * (double)kk / (double)N2;
        wN[kk] = make_float2((float)cos(ang), (float)sin(ang));
    }
    SlicqDeviceTables& d = p->dev;
    memset(&d, 0, sizeof d);
    d.L = L; d.N2 = p->N2; d.hop = p->hop; d.n_bins = J; d.n_buckets = (int)p->buckets.size(); d.sum_M = p->sum_M;
    d.pad_l = pad_l; d.pad_r = pad_r; d.tw_lo = tw_lo; d.tw_hi = tw_hi;
    d.adjoint = t->flags & (SLICQ_PLAN_ADJOINT_OF_SYNTHESIS | SLICQ_PLAN_ADJOINT_OF_ANALYSIS);
    d.spec_scale = (d.adjoint & 1) ? (float)(2.0 / L) : 1.f;
    d.ends_scale = (d.adjoint & 1) ? (float)(1.0 / L) : 1.f;
    int rc = 0;
    rc |= upload(tuk, &d.tukey, p->owned);
    rc |= upload(wf, &d.wf, p->owned);
    rc |= upload(wi, &d.wi, p->owned);
    rc |= upload(p->bin_pos, &d.bin_pos, p->owned);
    rc |= upload(p->bin_M, &d.bin_M, p->owned);
    rc |= upload(p->bin_coff, &d.bin_coff, p->owned);
    rc |= upload(post, &d.post_tw, p->owned);
    rc |= upload(tw, &d.tw, p->owned);
    rc |= upload(toff, &d.bin_toff, p->owned);
    rc |= upload(ov, &d.bin_ov, p->owned);
    rc |= upload(ovoff, &d.bin_ovoff, p->owned);
    rc |= upload(ex, &d.ex, p->owned);
    rc |= upload(gaps, &d.gaps, p->owned);
    rc |= upload(mir, &d.mir, p->owned);
    rc |= upload(wN, &d.wN, p->owned);
    rc |= upload(wP, &d.wP, p->owned);
    d.n_mir = n_mir; d.n_fac = (int)fac.size();
    for (size_t i = 0; i < fac.size(); ++i) d.fac[i] = fac[i];
    d.n_ex = n_ex; d.pl_off = pl_off; d.pl_len = pl_len; d.t_stride = (int)p->t_stride;
    if (rc) {
        slicq_plan_destroy(p);
        return fail(SLICQ_E_CUDA, "device table upload failed");
    }
    p->spec_stride_fwd = (pad_l + p->N2 + 1 + pad_r + 1) & ~1LL;  // even: 16-byte aligned rows
    const char* env = getenv("SLICQ_CHUNK_MB");
    long long mb = env ? atoll(env) : 2048;
    if (mb < 1) mb = 1;
    p->chunk_bytes = mb << 20;
    const char* envj = getenv("SLICQ_BINS_JOBS");
    // jobs per bins launch: 148 SMs x 3 resident CTAs x 10 waves for large launches (short jobs even out the tail;
    // swept 666 ... 8880 at batch 8 in both rounds), 1184 for small ones (a job keeps enough iterations to amortise its
    // set-up); SLICQ_BINS_JOBS fixes the number
    p->target_jobs = envj ? atoi(envj) : 0;
    if (p->target_jobs < 0) p->target_jobs = 0;
    const char* envs = getenv("SLICQ_SPLIT_UNITS");
    p->split_units = envs ? atoll(envs) : 1184;
    const char* envw = getenv("SLICQ_SPLIT_WAYS");
    p->split_ways = envw ? atoi(envw) : 2;
    if (p->split_ways < 2) p->split_ways = 2;
    if (p->split_ways > SLICQ_MAX_WAYS) p->split_ways = SLICQ_MAX_WAYS;
    p->side_ready = false;
    const char* envb = getenv("SLICQ_ONLY_BUCKET");
    p->only_bucket = envb ? atoi(envb) : -1;
    const char* envi = getenv("SLICQ_BINS_MIN_ITERS");
    p->min_iters = envi ? atoi(envi) : 1;
    if (p->min_iters < 1) p->min_iters = 1;
    *out = p;
    return SLICQ_OK;
}

extern "C" int slicq_plan_n_buckets(const slicq_plan* p) { return p ? (int)p->buckets.size() : SLICQ_E_INVALID; }

extern "C" int slicq_plan_bucket_info(const slicq_plan* p, int b, int32_t* first_bin, int32_t* n_bins, int32_t* M) {
    if (!p || b < 0 || b >= (int)p->buckets.size()) return fail(SLICQ_E_INVALID, "bad bucket index");
    if (first_bin) *first_bin = p->buckets[b].first_bin;
    if (n_bins) *n_bins = p->buckets[b].n_bins;
    if (M) *M = p->buckets[b].M;
    return SLICQ_OK;
}

extern "C" int64_t slicq_plan_num_slices(const slicq_plan* p, int64_t T) {
    if (!p || T < 0) return SLICQ_E_INVALID;
    const int64_t hh = p->L / 4;
    const int64_t nblk = (T + hh - 1) / hh;  // quarter-hop blocks (slicing.py:34-42)
    return (nblk + 1) / 2 + 1;               // slicing.py:49,61-72
}

namespace {
long long bytes_per_unit(const slicq_plan* p, int inverse) {
    if (!inverse) return p->spec_stride_fwd * 8;
    return p->t_stride * 8;
}
long long chunk_units(const slicq_plan* p, int inverse) {
    long long c = p->chunk_bytes / bytes_per_unit(p, inverse);
    if (c >= 296) c -= c % 296;   // whole waves of the slice kernels (148 SMs x 2 CTAs)
    return c < 1 ? 1 : c;
}
}  // namespace

namespace {
bool use_split(const slicq_plan* p, int64_t n_rows, int64_t n_slices) {
    return p->split_units > 0 && n_rows >= 2 && n_rows * n_slices >= p->split_units;
}
// rows are dealt to `ways` groups of (almost) equal size; group h = rows [part_row0(h), part_row0(h + 1))
int split_ways(const slicq_plan* p, int64_t n_rows) { return (int)(n_rows < p->split_ways ? n_rows : p->split_ways); }
int64_t part_row0(int64_t n_rows, int ways, int h) { return n_rows * h / ways; }
size_t half_scratch(const slicq_plan* p, int64_t rows, int64_t n_slices, int inverse) {
    long long units = rows * n_slices;
    const long long c = chunk_units(p, inverse);
    if (units > c) units = c;
    return ((size_t)(units * bytes_per_unit(p, inverse)) + 511) & ~(size_t)255;
}
int ensure_side_streams(const slicq_plan* p) {
    if (p->side_ready) return 0;
    if (cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming)) return -1;
    for (int h = 0; h < SLICQ_MAX_WAYS; ++h)
        if (cudaStreamCreateWithFlags(&p->side[h], cudaStreamNonBlocking) || cudaEventCreateWithFlags(&p->ev_join[h], cudaEventDisableTiming))
            return -1;
    p->side_ready = true;
    return 0;
}
}  // namespace

extern "C" size_t slicq_scratch_bytes(const slicq_plan* p, int64_t n_rows, int64_t n_slices, int inverse) {
    if (!p || n_rows <= 0 || n_slices <= 0) return 0;
    if (use_split(p, n_rows, n_slices)) {
        const int ways = split_ways(p, n_rows);
        size_t total = 256;
        for (int h = 0; h < ways; ++h)
            total += half_scratch(p, part_row0(n_rows, ways, h + 1) - part_row0(n_rows, ways, h), n_slices, inverse);
        return total;
    }
    return half_scratch(p, n_rows, n_slices, inverse) + 256;
}

namespace {
int fill_bins_params(const slicq_plan* p, const slicq_bucket_view* views, SlicqBinsParams& bp, int n_rs,
                     const slicq_bucket_view* masks = nullptr, const slicq_bucket_view* norms = nullptr) {
    int jobs = 0;
    bp.n_buckets = (int)p->buckets.size();
    double total = 0.0;
    for (const Bucket& b : p->buckets) total += b.cost;
    const int target = p->target_jobs > 0 ? p->target_jobs : std::min(4440, std::max(1184, 4 * n_rs));
    for (size_t i = 0; i < p->buckets.size(); ++i) {
        const Bucket& b = p->buckets[i];
        SlicqBucketArg& a = bp.b[i];
        a.ptr = reinterpret_cast<float2*>(views[i].ptr);
        a.s_row = views[i].s_row; a.s_bin = views[i].s_bin; a.s_slice = views[i].s_slice;
        a.M = b.M; a.first_bin = b.first_bin; a.n_bins = b.n_bins; a.gt = b.gt; a.tw_off = b.tw_off;
        a.gap_first = b.gap_first; a.gap_n = b.gap_n;
        a.mptr = masks ? reinterpret_cast<const float*>(masks[i].ptr) : nullptr;
        a.nptr = norms ? reinterpret_cast<float*>(norms[i].ptr) : nullptr;
        const slicq_bucket_view* aux = masks ? masks : norms;      // synthesis masks or analysis magnitudes: never both
        a.ms_row = aux ? aux[i].s_row : 0; a.ms_bin = aux ? aux[i].s_bin : 0; a.ms_slice = aux ? aux[i].s_slice : 0;
        if (p->only_bucket >= 0 && (int)i != p->only_bucket) {       // tuning aid: time one bucket alone
            a.units_per_job = b.gt; a.n_jobs = 0; a.job_start = jobs;
            continue;
        }
        const int groups = (n_rs + b.gt - 1) / b.gt;                 // iterations available in this chunk
        int nj = (int)(target * b.cost / total + 0.5);
        if (nj > groups / p->min_iters) nj = groups / p->min_iters;
        if (nj < 1) nj = 1;
        const int gpj = (groups + nj - 1) / nj;                      // iterations per job
        a.units_per_job = gpj * b.gt;
        a.n_jobs = (n_rs + a.units_per_job - 1) / a.units_per_job;
        a.job_start = jobs;
        jobs += a.n_jobs;
    }
    return jobs;
}
}  // namespace

namespace {
int forward_one(const slicq_plan* p, const float* x, int64_t n_rows, int64_t x_row_stride,
                int64_t n_samples, int64_t t0, int64_t k0, int64_t n_slices,
                const slicq_bucket_view* buckets, const slicq_bucket_view* norms, void* scratch, cudaStream_t s);
int inverse_one(const slicq_plan* p, const slicq_bucket_view* buckets, const slicq_bucket_view* masks, int64_t x_rows,
                int64_t n_rows, int64_t n_slices, int64_t k0, float* y, int64_t y_row_stride, int64_t length,
                int64_t t0, float* halo_out, void* scratch, cudaStream_t s);

// run `fn(row0, rows, scratch, stream)` for the row groups on the plan's side streams
template <class F>
int run_split(const slicq_plan* p, int64_t n_rows, int64_t n_slices, int inverse, void* scratch, cudaStream_t s, F fn) {
    std::lock_guard<std::mutex> lock(p->split_mutex);
    if (ensure_side_streams(p)) return fail(SLICQ_E_CUDA, "cannot create internal streams");
    const int ways = split_ways(p, n_rows);
    unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
    cudaEventRecord(p->ev_fork, s);
    int rc = 0;
    for (int h = 0; h < ways && rc == 0; ++h) {
        const int64_t r0 = part_row0(n_rows, ways, h), rows = part_row0(n_rows, ways, h + 1) - r0;
        cudaStreamWaitEvent(p->side[h], p->ev_fork, 0);
        rc = fn(r0, rows, base, p->side[h]);
        base += half_scratch(p, rows, n_slices, inverse);
        cudaEventRecord(p->ev_join[h], p->side[h]);
        cudaStreamWaitEvent(s, p->ev_join[h], 0);
    }
    return rc;
}
}  // namespace

namespace {
int forward_impl(const slicq_plan* p, const float* x, int64_t n_rows, int64_t x_row_stride,
                 int64_t n_samples, int64_t t0, int64_t k0, int64_t n_slices,
                 const slicq_bucket_view* buckets, const slicq_bucket_view* norms, void* scratch, size_t scratch_bytes,
                 void* stream) {
    if (!p || !x || !buckets) return fail(SLICQ_E_INVALID, "null argument");
    if (n_rows <= 0 || n_slices <= 0 || n_samples < 0) return fail(SLICQ_E_INVALID, "bad shape");
    if (n_rows * n_slices > 0x7fffffffLL) return fail(SLICQ_E_INVALID, "too many (row,slice) units");
    if (scratch_bytes < slicq_scratch_bytes(p, n_rows, n_slices, 0) || !scratch)
        return fail(SLICQ_E_SCRATCH, "scratch buffer too small (see slicq_scratch_bytes)");
    if (reinterpret_cast<uintptr_t>(scratch) & 15) return fail(SLICQ_E_SCRATCH, "scratch buffer must be 16-byte aligned");
    for (size_t i = 0; i < p->buckets.size(); ++i)
        if (!buckets[i].ptr || (norms && !norms[i].ptr)) return fail(SLICQ_E_INVALID, "null bucket pointer");
    cudaStream_t s0 = reinterpret_cast<cudaStream_t>(stream);
    if (use_split(p, n_rows, n_slices)) {
        return run_split(p, n_rows, n_slices, 0, scratch, s0, [&](int64_t r0, int64_t rows, void* scr, cudaStream_t st) {
            std::vector<slicq_bucket_view> v(buckets, buckets + p->buckets.size()), nv;
            for (auto& b : v) b.ptr = reinterpret_cast<float2*>(b.ptr) + r0 * b.s_row;
            if (norms) {
                nv.assign(norms, norms + p->buckets.size());
                for (auto& b : nv) b.ptr = reinterpret_cast<float*>(b.ptr) + r0 * b.s_row;
            }
            return forward_one(p, x + r0 * x_row_stride, rows, x_row_stride, n_samples, t0, k0, n_slices, v.data(),
                               norms ? nv.data() : nullptr, scr, st);
        });
    }
    return forward_one(p, x, n_rows, x_row_stride, n_samples, t0, k0, n_slices, buckets, norms, scratch, s0);
}
}  // namespace

extern "C" int slicq_forward(const slicq_plan* p, const float* x, int64_t n_rows, int64_t x_row_stride,
                             int64_t n_samples, int64_t t0, int64_t k0, int64_t n_slices,
                             const slicq_bucket_view* buckets, void* scratch, size_t scratch_bytes, void* stream) {
    return forward_impl(p, x, n_rows, x_row_stride, n_samples, t0, k0, n_slices, buckets, nullptr, scratch, scratch_bytes, stream);
}

extern "C" int slicq_forward_norm(const slicq_plan* p, const float* x, int64_t n_rows, int64_t x_row_stride,
                                  int64_t n_samples, int64_t t0, int64_t k0, int64_t n_slices,
                                  const slicq_bucket_view* buckets, const slicq_bucket_view* norms, void* scratch,
                                  size_t scratch_bytes, void* stream) {
    if (!norms) return fail(SLICQ_E_INVALID, "norms missing");
    return forward_impl(p, x, n_rows, x_row_stride, n_samples, t0, k0, n_slices, buckets, norms, scratch, scratch_bytes, stream);
}

namespace {
int forward_one(const slicq_plan* p, const float* x, int64_t n_rows, int64_t x_row_stride,
                int64_t n_samples, int64_t t0, int64_t k0, int64_t n_slices,
                const slicq_bucket_view* buckets, const slicq_bucket_view* norms, void* scratch, cudaStream_t s) {
    const long long units = n_rows * n_slices, cu = chunk_units(p, 0);
    float2* H = reinterpret_cast<float2*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
    SlicqSliceParams sp;
    memset(&sp, 0, sizeof sp);
    sp.t = p->dev; sp.x = const_cast<float*>(x); sp.x_row_stride = x_row_stride; sp.T = n_samples; sp.t0 = t0; sp.k0 = k0;
    sp.spec = H; sp.spec_stride = p->spec_stride_fwd; sp.S = (int)n_slices;
    SlicqBinsParams* bp = new SlicqBinsParams();
    bp->t = p->dev; bp->spec = H; bp->spec_stride = p->spec_stride_fwd; bp->S = (int)n_slices;
    int rc = 0;
    for (long long u0 = 0; u0 < units && rc == 0; u0 += cu) {
        const int n = (int)((units - u0 < cu) ? (units - u0) : cu);
        sp.n_rs = n; sp.rs0 = (int)u0;
        { ProfScope ps(K_SLICE_FWD, s); rc = slicq_launch_slice_fwd(&sp, s); }
        ++g_launches;
        if (rc) break;
        bp->n_rs = n; bp->rs0 = (int)u0;
        const int jobs = fill_bins_params(p, buckets, *bp, n, nullptr, norms);
        { ProfScope ps(K_BINS_FWD, s); rc = slicq_launch_bins(bp, jobs, p->bins_smem, 0, s); }
        ++g_launches;
    }
    delete bp;
    if (rc) {
        char buf[128];
        snprintf(buf, sizeof buf, "kernel launch failed: %s", rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "unsupported");
        return fail(SLICQ_E_CUDA, buf);
    }
    return SLICQ_OK;
}
}  // namespace

// Canonical packed layout: all buckets in one allocation, bucket b = contiguous [n_rows][F_b][S][M_b].
namespace {
// canonical packed layout: bucket b = contiguous [n_rows][F_b][S][M_b] elements of `elem` bytes, buckets in order
std::vector<slicq_bucket_view> packed_views(const slicq_plan* p, void* base_, int64_t n_rows, int64_t n_slices, size_t elem) {
    std::vector<slicq_bucket_view> v(p->buckets.size());
    unsigned char* base = reinterpret_cast<unsigned char*>(base_);
    for (size_t i = 0; i < p->buckets.size(); ++i) {
        const Bucket& b = p->buckets[i];
        v[i].ptr = base;
        v[i].s_slice = b.M;
        v[i].s_bin = (int64_t)n_slices * b.M;
        v[i].s_row = (int64_t)b.n_bins * n_slices * b.M;
        base += (size_t)n_rows * b.n_bins * n_slices * b.M * elem;
    }
    return v;
}
}  // namespace

// Canonical packed layout: all buckets in one allocation, bucket b = contiguous [n_rows][F_b][S][M_b].
extern "C" int slicq_forward_packed(const slicq_plan* p, const float* x, int64_t n_rows, int64_t x_row_stride,
                                    int64_t n_samples, int64_t t0, int64_t k0, int64_t n_slices, void* coefs,
                                    void* scratch, size_t scratch_bytes, void* stream) {
    if (!p || !coefs) return fail(SLICQ_E_INVALID, "null argument");
    std::vector<slicq_bucket_view> v = packed_views(p, coefs, n_rows, n_slices, 8);
    return slicq_forward(p, x, n_rows, x_row_stride, n_samples, t0, k0, n_slices, v.data(), scratch, scratch_bytes, stream);
}

extern "C" int slicq_forward_packed_norm(const slicq_plan* p, const float* x, int64_t n_rows, int64_t x_row_stride,
                                         int64_t n_samples, int64_t t0, int64_t k0, int64_t n_slices, void* coefs,
                                         void* norms, void* scratch, size_t scratch_bytes, void* stream) {
    if (!p || !coefs || !norms) return fail(SLICQ_E_INVALID, "null argument");
    std::vector<slicq_bucket_view> v = packed_views(p, coefs, n_rows, n_slices, 8);
    std::vector<slicq_bucket_view> nv = packed_views(p, norms, n_rows, n_slices, 4);
    return slicq_forward_norm(p, x, n_rows, x_row_stride, n_samples, t0, k0, n_slices, v.data(), nv.data(), scratch,
                              scratch_bytes, stream);
}

namespace {
// n_rows = output rows; masks == nullptr: plain synthesis of `buckets` (n_rows rows);
// else buckets hold the mixture (x_rows rows) and masks the per-output-row fp32 masks.
int inverse_impl(const slicq_plan* p, const slicq_bucket_view* buckets, const slicq_bucket_view* masks, int64_t x_rows,
                 int64_t n_rows, int64_t n_slices, int64_t k0, float* y, int64_t y_row_stride, int64_t length,
                 int64_t t0, float* halo_out, void* scratch, size_t scratch_bytes, void* stream) {
    // y may be null when length == 0 (a shard that owns no output samples still produces its halo)
    if (!p || (!y && length > 0) || !buckets) return fail(SLICQ_E_INVALID, "null argument");
    if (n_rows <= 0 || n_slices <= 0 || length < 0) return fail(SLICQ_E_INVALID, "bad shape");
    if (n_rows * n_slices > 0x7fffffffLL) return fail(SLICQ_E_INVALID, "too many (row,slice) units");
    if (scratch_bytes < slicq_scratch_bytes(p, n_rows, n_slices, 1) || !scratch)
        return fail(SLICQ_E_SCRATCH, "scratch buffer too small (see slicq_scratch_bytes)");
    if (reinterpret_cast<uintptr_t>(scratch) & 15) return fail(SLICQ_E_SCRATCH, "scratch buffer must be 16-byte aligned");
    for (size_t i = 0; i < p->buckets.size(); ++i)
        if (!buckets[i].ptr) return fail(SLICQ_E_INVALID, "null bucket pointer");
    cudaStream_t s0 = reinterpret_cast<cudaStream_t>(stream);
    // masked synthesis: output row r reads mixture row r % x_rows, so a row group must start on a multiple of x_rows
    bool split = use_split(p, n_rows, n_slices);
    if (split && masks) {
        const int ways = split_ways(p, n_rows);
        for (int h = 1; h < ways; ++h) split = split && (part_row0(n_rows, ways, h) % x_rows == 0);
    }
    if (split) {
        return run_split(p, n_rows, n_slices, 1, scratch, s0, [&](int64_t r0, int64_t rows, void* scr, cudaStream_t st) {
            std::vector<slicq_bucket_view> v(buckets, buckets + p->buckets.size()), mv;
            if (masks) {        // the mixture is shared by all row groups, the masks follow the output rows
                mv.assign(masks, masks + p->buckets.size());
                for (auto& b : mv) b.ptr = reinterpret_cast<float*>(b.ptr) + r0 * b.s_row;
            } else {
                for (auto& b : v) b.ptr = reinterpret_cast<float2*>(b.ptr) + r0 * b.s_row;
            }
            return inverse_one(p, v.data(), masks ? mv.data() : nullptr, x_rows, rows, n_slices, k0, y + r0 * y_row_stride,
                               y_row_stride, length, t0, halo_out ? halo_out + r0 * p->hop : nullptr, scr, st);
        });
    }
    return inverse_one(p, buckets, masks, x_rows, n_rows, n_slices, k0, y, y_row_stride, length, t0, halo_out, scratch, s0);
}

int inverse_one(const slicq_plan* p, const slicq_bucket_view* buckets, const slicq_bucket_view* masks, int64_t x_rows,
                int64_t n_rows, int64_t n_slices, int64_t k0, float* y, int64_t y_row_stride, int64_t length,
                int64_t t0, float* halo_out, void* scratch, cudaStream_t s) {
    const long long units = n_rows * n_slices, cu = chunk_units(p, 1);
    const long long nu = units < cu ? units : cu;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
    float2* T = reinterpret_cast<float2*>(base);
    (void)nu;
    SlicqBinsParams* bp = new SlicqBinsParams();
    bp->t = p->dev; bp->spec = T; bp->spec_stride = p->t_stride; bp->S = (int)n_slices;
    bp->x_rows = masks ? (int)x_rows : 0;
    bp->k0 = (int)k0;
    SlicqSliceParams sp;
    memset(&sp, 0, sizeof sp);
    sp.t = p->dev; sp.k0 = k0; sp.spec = T; sp.spec_stride = p->t_stride; sp.S = (int)n_slices;
    sp.x = y; sp.x_row_stride = y_row_stride; sp.T = length; sp.t0 = t0; sp.halo_out = halo_out;
    int rc = 0;
    for (long long u0 = 0; u0 < units && rc == 0;) {
        long long n = (units - u0 < cu) ? (units - u0) : cu;
        // a chunk must not start on an even slice k > 0: the odd slice before it accumulates into
        // hops that the even slice has to have stored first (see slice_fft_inv_kernel)
        while (n > 1 && u0 + n < units && ((u0 + n) % n_slices) != 0 && (((u0 + n) % n_slices) & 1) == 0) --n;
        bp->n_rs = (int)n; bp->rs0 = (int)u0;
        const int jobs = fill_bins_params(p, buckets, *bp, (int)n, masks);
        { ProfScope ps(K_BINS_INV, s); rc = slicq_launch_bins(bp, jobs, p->bins_smem, 1, s); }
        ++g_launches;
        if (rc) break;
        if (p->dev.n_mir > 0) {          // mirrored-bin pass of configurations whose first bins reach below DC
            rc = slicq_launch_mirror_fix(bp, s);
            ++g_launches;
            if (rc) break;
        }
        sp.n_rs = (int)n; sp.rs0 = (int)u0;
        {
            ProfScope ps(K_SLICE_INV, s);
            sp.parity = 0; rc = slicq_launch_slice_inv(&sp, s);
            if (!rc) { sp.parity = 1; rc = slicq_launch_slice_inv(&sp, s); }
        }
        g_launches += 2;
        u0 += n;
    }
    delete bp;
    if (rc) {
        char buf[128];
        snprintf(buf, sizeof buf, "kernel launch failed: %s", rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "unsupported");
        return fail(SLICQ_E_CUDA, buf);
    }
    return SLICQ_OK;
}
}  // namespace

extern "C" int slicq_inverse(const slicq_plan* p, const slicq_bucket_view* buckets, int64_t n_rows,
                             int64_t n_slices, int64_t k0, float* y, int64_t y_row_stride, int64_t length,
                             int64_t t0, float* halo_out, void* scratch, size_t scratch_bytes, void* stream) {
    return inverse_impl(p, buckets, nullptr, 0, n_rows, n_slices, k0, y, y_row_stride, length, t0, halo_out,
                        scratch, scratch_bytes, stream);
}

extern "C" int slicq_inverse_masked(const slicq_plan* p, const slicq_bucket_view* mix, const slicq_bucket_view* masks,
                                    int64_t n_targets, int64_t n_rows, int64_t n_slices, int64_t k0, float* y,
                                    int64_t y_row_stride, int64_t length, int64_t t0, float* halo_out, void* scratch,
                                    size_t scratch_bytes, void* stream) {
    if (!masks || n_targets <= 0) return fail(SLICQ_E_INVALID, "masks / n_targets missing");
    if (p) for (size_t i = 0; i < p->buckets.size(); ++i)
        if (!masks[i].ptr) return fail(SLICQ_E_INVALID, "null mask pointer");
    return inverse_impl(p, mix, masks, n_rows, n_targets * n_rows, n_slices, k0, y, y_row_stride, length, t0, halo_out,
                        scratch, scratch_bytes, stream);
}

