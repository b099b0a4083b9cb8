// rr_abi.cu -- the C ABI of include/rr_maxcorr.h: device packing, scan orchestration,
// multi-GPU partition and merge.  Host logic here is O(R log R + N); everything per pair
// runs in the CUDA kernels.  There is no CPU implementation of the scan in this library.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "rr_kernels.h"
#include "rr_device.cuh"
#include "rr_plan.h"
#include "rr_kmeans.h"
#include "../../include/rr_debug.h"

#include <chrono>
static double rr_now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define RR_TRACE(tag) rr_trace_mark(tag)
static bool rr_trace_on() { return getenv("RR_TRACE") != nullptr; }

#define RR_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            rr_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
            return RR_E_CUDA;                                                                      \
        }                                                                                          \
    } while (0)

// ---------------------------------------------------------------------------------------
// host buffers
// ---------------------------------------------------------------------------------------
extern "C" int rr_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int rr_variant_available(int variant)
{
    if (variant == RR_VARIANT_BITSET || variant == RR_VARIANT_AUTO) return 1;
    if (variant == RR_VARIANT_UMMA || variant == RR_VARIANT_UMMA_F4 || variant == RR_VARIANT_UMMA_MXF4) return rr_umma_available();
    return 0;
}

extern "C" void *rr_host_alloc(size_t bytes, int *pinned)
{
    void *p = nullptr;
    *pinned = 0;
    if (rr_device_count() > 0 && cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) {
        *pinned = 1;
        return p;
    }
    cudaGetLastError();
    return malloc(bytes);
}

extern "C" void rr_host_free(void *p, int pinned)
{
    if (!p) return;
    if (pinned) cudaFreeHost(p);
    else free(p);
}

// first CUDA call of a process = context creation (a few hundred ms): rr_msa_read starts it on a background thread
// while the text is being indexed
static std::mutex g_warm_mu;
static std::thread g_warm_thread;
extern "C" void rr_cuda_warmup_end(void)
{
    std::lock_guard<std::mutex> lk(g_warm_mu);
    if (g_warm_thread.joinable()) g_warm_thread.join();
}
extern "C" void rr_cuda_warmup_begin(void)
{
    static bool once = false;
    std::lock_guard<std::mutex> lk(g_warm_mu);
    if (once) return;
    once = true;
    g_warm_thread = std::thread([] {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return; }
        if (cudaSetDevice(0) == cudaSuccess) cudaFree(nullptr);
        cudaGetLastError();
    });
    atexit(rr_cuda_warmup_end);
}

// ---------------------------------------------------------------------------------------
// device memory: stream-ordered pool
// ---------------------------------------------------------------------------------------
static thread_local cudaStream_t g_alloc_stream = nullptr;
void rr_alloc_stream(cudaStream_t st) { g_alloc_stream = st; }

cudaError_t rr_dev_malloc(void **p, size_t bytes)
{
    static std::atomic<unsigned> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(configured.load() & (1u << dev))) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        configured |= (1u << dev);
    }
    return cudaMallocAsync(p, bytes, g_alloc_stream);
}

void rr_dev_free(void *p)
{
    if (p) cudaFreeAsync(p, g_alloc_stream);
}

// The length classes of the row order (rr_plan.h): rows sorted by span length are cut at fixed fractions, rounded down to
// whole 256-row K blocks; below 1024 rows everything is one class (the last).  class_start[RR_LENGTH_CLASSES + 1].
static const double RR_CLASS_FRACTION[RR_LENGTH_CLASSES] = RR_LENGTH_CLASS_FRACTIONS;
extern "C" int rr_length_classes(int rows, int32_t *class_start)
{
    if (rows < 0 || !class_start) { rr_set_error("rr_length_classes: bad arguments"); return -1; }
    class_start[0] = 0;
    double f = 0.0;
    for (int c = 1; c < RR_LENGTH_CLASSES; c++) {
        f += RR_CLASS_FRACTION[c - 1];
        class_start[c] = rows >= 1024 ? (int)((int64_t)((double)rows * f) / 256 * 256) : 0;
        if (class_start[c] < class_start[c - 1]) class_start[c] = class_start[c - 1];
    }
    class_start[RR_LENGTH_CLASSES] = rows;
    return RR_LENGTH_CLASSES;
}

extern "C" int rr_contraction_ranges(const int32_t *start, const int32_t *end, int rows, int cols, int n_classes,
                                     const int32_t *class_start, int ti, int tj, int kunit, int32_t *k_lo, int32_t *k_hi,
                                     int *n_rowblocks)
{
    if (rows < 0 || cols < 0 || ti < 1 || tj < 1 || kunit < 1 || n_classes < 1 || n_classes > RR_MAX_CLASSES || !class_start ||
        !k_lo || !k_hi || !n_rowblocks || (rows && (!start || !end))) {
        rr_set_error("rr_contraction_ranges: bad arguments");
        return RR_E_ARG;
    }
    const int ncb = std::max((cols + tj - 1) / tj, 1);
    if (rows < 4) {   // too few rows for an admissible group (size < R): no row site, nothing contributes
        std::fill(k_lo, k_lo + (size_t)n_classes * ncb, 0);
        std::fill(k_hi, k_hi + n_classes, 0);
        *n_rowblocks = 0;
        return RR_OK;
    }
    // every group admissible as a row and as a column group, no first-break: all sites with a partner are row sites
    const int mincov = 4;
    std::vector<int32_t> gsize((size_t)5 * cols), coverage(cols, 20), breakcol(cols, cols);
    for (size_t g = 0; g < gsize.size(); g++) gsize[g] = (g % 5 < 4) ? 3 : 2;   // mincov/4 < size < rows; 4 x 3 bases > 20 / 2
    rr_plan plan;
    rr_plan_build(plan, rows, cols, mincov, gsize.data(), coverage.data(), breakcol.data(), start, end, class_start, n_classes,
                  ti, tj, kunit, 1, 0, 0, 1);
    if ((int)plan.k_lo.size() != n_classes * std::max(plan.n_colblocks, 1) || (int)plan.k_hi.size() != n_classes * std::max(plan.n_rowblocks, 1)) {
        rr_set_error("rr_contraction_ranges: unexpected plan shape");
        return RR_E_ARG;
    }
    std::copy(plan.k_lo.begin(), plan.k_lo.end(), k_lo);
    std::copy(plan.k_hi.begin(), plan.k_hi.end(), k_hi);
    *n_rowblocks = plan.n_rowblocks;
    return RR_OK;
}

// ---------------------------------------------------------------------------------------
// measurement helpers
// ---------------------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
extern "C" void rr_count_launch(int n) { g_launches += n; }
extern "C" int64_t rr_launch_count(void) { return (int64_t)g_launches.load(); }

// ---------------------------------------------------------------------------------------
// host builds of the score functions (tests, RR_FLAG_HOST_FINALIZE)
// ---------------------------------------------------------------------------------------
static std::vector<double> &host_lnfact(size_t need)
{
    static thread_local std::vector<double> tab;
    if (tab.size() < need) {
        size_t old = tab.size();
        tab.resize(std::max(need, old * 2 + 1024));
        for (size_t n = old; n < tab.size(); n++) tab[n] = rr_lnfact((unsigned)n);
    }
    return tab;
}

extern "C" double rr_score_host(uint32_t s, uint32_t gr1, uint32_t gr2, uint32_t cov, int32_t sizei, int32_t sizej)
{
    std::vector<double> &t = host_lnfact((size_t)cov + 2);
    return rr_positive_significance(t.data(), s, gr1, gr2, cov, sizei, sizej);
}

extern "C" double rr_score_bound_host(uint32_t s, uint32_t gr1, uint32_t gr2, uint32_t cov)
{
    std::vector<double> &t = host_lnfact((size_t)cov + 2);
    return rr_bound_effective(rr_score_upper_bound(t.data(), s, gr1, gr2, cov));
}

extern "C" int rr_below_median_host(uint32_t s, uint32_t gr1, uint32_t gr2, uint32_t cov)
{
    return rr_below_median(s, gr1, gr2, cov);
}

// ---------------------------------------------------------------------------------------
// packed MSA
// ---------------------------------------------------------------------------------------
struct scan_buffers {
    uint8_t *rowok = nullptr, *colok = nullptr;
    int32_t *breakcol = nullptr, *rowsites = nullptr, *unit_cb0 = nullptr, *word_hi = nullptr, *word_lo = nullptr;
    int64_t *unit_prefix = nullptr;
    void release()
    {
        rr_dev_free(rowok); rr_dev_free(colok); rr_dev_free(breakcol); rr_dev_free(rowsites); rr_dev_free(unit_cb0);
        rr_dev_free(word_hi); rr_dev_free(word_lo); rr_dev_free(unit_prefix);
        rowok = colok = nullptr; breakcol = rowsites = unit_cb0 = word_hi = word_lo = nullptr; unit_prefix = nullptr;
    }
    ~scan_buffers() { release(); }
};

// the host plan and its device copies are kept between scans of the same (mincov, variant, part, break mode)
struct scan_cache {
    bool valid = false;
    int mincov = 0, variant = 0, part_index = 0, part_count = 0;
    bool general = false;
    uint64_t plan_id = 0;
    rr_plan plan;
    scan_buffers sb;
};

struct rr_packed {
    int device = 0, n_sm = 148;
    int R = 0, N = 0, W32 = 0, codes = 0;
    int phase = 0;                        // 1: rows uploaded, 2: bitsets packed (this slice), 3: complete
    int row_lo = 0, row_hi = 0;           // the rows this handle uploaded (rr_pack_rows)
    std::vector<int32_t> slice_spans;     // [3][row_hi - row_lo] start, end, covered cells
    cudaEvent_t pe0 = nullptr, pe1 = nullptr, pe2 = nullptr;
    cudaStream_t st = nullptr;
    uint8_t *d_cells = nullptr;
    int32_t *d_perm = nullptr;
    uint32_t *d_bits = nullptr, *d_covbits = nullptr;   // one allocation: [5N][W32] then [N][W32]
    int32_t *d_gsize = nullptr, *d_coverage = nullptr;
    double *d_lnfact = nullptr;
    rr_best_t *d_best = nullptr;
    unsigned long long *d_counters = nullptr;
    rr_cand_p *d_deferred = nullptr;      // candidates of the tcgen05 scan awaiting their exact evaluation (rr_device.cuh)
    size_t deferred_cap = 0;
    std::vector<int32_t> h_start, h_end;  // spans in rank order: (length class, span start, span end)
    std::vector<int32_t> h_perm;          // rank -> row of the MSA
    int32_t class_start[RR_MAX_CLASSES + 1] = {0};   // rank boundaries of the length classes (rr_length_classes)
    int n_classes = 1;
    std::vector<int32_t> h_gsize, h_coverage;
    bool contiguous = true;
    float h2d_ms = 0.f, pack_ms = 0.f;
    bool have_result = false;
    rr_umma_state *umma = nullptr;  // int8 operands + tensor maps, built on first use
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    scan_cache cache;
    uint64_t next_plan_id = 1;
};

template <typename T>
static int dev_alloc(T **p, size_t count)
{
    *p = nullptr;
    if (rr_dev_malloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        rr_set_error("out of device memory (%zu bytes)", count * sizeof(T));
        return RR_E_NOMEM;
    }
    return RR_OK;
}

// device buffers of one call, returned to the pool on every exit path
struct dev_scope {
    std::vector<void *> ptrs;
    template <typename T>
    int alloc(T **p, size_t count)
    {
        int rc = dev_alloc(p, count);
        if (!rc) ptrs.push_back(*p);
        return rc;
    }
    ~dev_scope() { for (void *p : ptrs) rr_dev_free(p); }
};

struct event_scope {
    std::vector<cudaEvent_t> evs;
    int create(cudaEvent_t *e)
    {
        if (cudaEventCreate(e) != cudaSuccess) { cudaGetLastError(); return 1; }
        evs.push_back(*e);
        return 0;
    }
    ~event_scope() { for (cudaEvent_t e : evs) cudaEventDestroy(e); }
};

extern "C" void rr_packed_free(rr_packed *pk)
{
    if (!pk) return;
    cudaSetDevice(pk->device);
    rr_alloc_stream(pk->st);
    rr_umma_free(pk->umma);
    rr_dev_free(pk->d_cells); rr_dev_free(pk->d_perm); rr_dev_free(pk->d_bits);
    rr_dev_free(pk->d_gsize); rr_dev_free(pk->d_coverage); rr_dev_free(pk->d_lnfact); rr_dev_free(pk->d_best);
    rr_dev_free(pk->d_counters); rr_dev_free(pk->d_deferred);
    pk->cache.sb.release();  // while the stream still exists
    if (pk->t0) { cudaEventDestroy(pk->t0); cudaEventDestroy(pk->t1); }
    if (pk->pe0) cudaEventDestroy(pk->pe0);
    if (pk->pe1) cudaEventDestroy(pk->pe1);
    if (pk->pe2) cudaEventDestroy(pk->pe2);
    if (pk->st) { cudaStreamSynchronize(pk->st); cudaStreamDestroy(pk->st); }
    rr_alloc_stream(nullptr);
    delete pk;
    RR_TRACE("free: done");
}

extern "C" void rr_msa_gather_rows(const rr_msa *m, int r0, int r1, uint8_t *dst, int threads);

// Rows of a text-backed or pageable host MSA -> d_cells through a small page-locked ring: a few host threads gather
// whole rows into ring slots (this is the only host copy the rows ever see) while earlier slots are in flight on the
// copy engine.  Page-locking the whole 1-2 GB matrix instead costs ~1 s per GB.
static int staged_upload(const rr_msa *msa, int row_lo, int row_hi, uint8_t *d_cells, int device, cudaStream_t st)
{
    const int R = row_hi - row_lo;                                       // rows [row_lo, row_hi) -> d_cells[0 .. R)
    const size_t N = (size_t)msa->cols;
    constexpr int NSLOT = 6, NTHREAD = 6;
    constexpr size_t SLOT_BYTES = (size_t)8 << 20;
    const int rows_per_slot = (int)std::max<size_t>(1, SLOT_BYTES / N);
    const size_t slot_bytes = (size_t)rows_per_slot * N;
    const int n_chunks = (R + rows_per_slot - 1) / rows_per_slot;
    uint8_t *ring = nullptr;
    RR_CUDA(cudaHostAlloc((void **)&ring, slot_bytes * NSLOT, cudaHostAllocDefault));
    RR_TRACE("pack: ring page-locked");
    cudaStream_t cst;
    cudaEvent_t ev[NSLOT];
    cudaError_t err0 = cudaStreamCreateWithFlags(&cst, cudaStreamNonBlocking);
    for (int k = 0; k < NSLOT; k++) if (err0 == cudaSuccess) err0 = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
    if (err0 != cudaSuccess) { cudaFreeHost(ring); rr_set_error("CUDA error %s (staged upload set-up)", cudaGetErrorString(err0)); return RR_E_CUDA; }
    std::mutex mu;
    std::condition_variable cv;
    int recorded[NSLOT] = {0};          // how many uses of the slot have had their copy issued + event recorded
    std::atomic<int> next{0};
    std::atomic<int> failed{0};
    auto work = [&]() {
        cudaSetDevice(device);
        for (;;) {
            const int c = next.fetch_add(1);
            if (c >= n_chunks || failed.load()) return;
            const int slot = c % NSLOT, use = c / NSLOT;
            {   // the previous use of this slot must have been issued, then completed
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return recorded[slot] == use || failed.load(); });
            }
            if (failed.load()) return;
            if (use > 0 && cudaEventSynchronize(ev[slot]) != cudaSuccess) { failed = 1; cv.notify_all(); return; }
            const int r0 = c * rows_per_slot, r1 = std::min(R, r0 + rows_per_slot);
            uint8_t *buf = ring + (size_t)slot * slot_bytes;
            rr_msa_gather_rows(msa, row_lo + r0, row_lo + r1, buf, 1);
            cudaError_t e = cudaMemcpyAsync(d_cells + (size_t)r0 * N, buf, (size_t)(r1 - r0) * N, cudaMemcpyHostToDevice, cst);
            {
                std::lock_guard<std::mutex> lk(mu);   // record under the lock: events of one stream stay in issue order
                if (e == cudaSuccess) e = cudaEventRecord(ev[slot], cst);
                if (e != cudaSuccess) failed = 1;
                recorded[slot] = use + 1;
            }
            cv.notify_all();
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < std::min(NTHREAD, n_chunks); t++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
    }
    cudaError_t e = failed.load() ? cudaErrorUnknown : cudaSuccess;
    cudaEvent_t done;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
    if (e == cudaSuccess) {
        e = cudaEventRecord(done, cst);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, done, 0);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cst);   // the ring is freed below
        cudaEventDestroy(done);
    }
    for (int k = 0; k < NSLOT; k++) cudaEventDestroy(ev[k]);
    cudaStreamDestroy(cst);
    cudaFreeHost(ring);
    if (e != cudaSuccess) { cudaGetLastError(); rr_set_error("CUDA error during the staged upload of the MSA (%s)", cudaGetErrorString(e)); return RR_E_CUDA; }
    return RR_OK;
}

// Packing runs in three phases so that several GPUs can share the work (rr_pack_rows / rr_pack_set_spans / rr_pack_finish,
// include/rr_maxcorr.h); rr_pack is the three in a row for the whole MSA on one GPU.
//   A  upload the rows [row_lo, row_hi) and find their covered spans
//   B  given the spans of ALL rows: row order (length class, span start, span end), then the group / coverage bitsets of
//      the uploaded rows at their ranks, zero elsewhere - the buffers of all slices OR (= add) to the full bitsets
//   C  group sizes, coverage, ln n! table, result buffers
static int pack_phase_a(const rr_msa *msa, int device, int row_lo, int row_hi, rr_packed *pk)
{
    const uint8_t *cells = msa->cells;
    const int R = msa->rows, N = msa->cols, codes = msa->codes;
    if (row_lo < 0 || row_hi > R || row_lo > row_hi) { rr_set_error("rr_pack_rows: rows [%d, %d) of %d", row_lo, row_hi, R); return RR_E_ARG; }
    rr_cuda_warmup_end();
    RR_TRACE("pack: context ready");
    int ndev = rr_device_count();
    if (ndev <= 0) { rr_set_error("no CUDA device: the scan has no CPU fallback"); return RR_E_NODEV; }
    if (device < 0 || device >= ndev) { rr_set_error("device %d out of range (%d devices)", device, ndev); return RR_E_ARG; }
    RR_CUDA(cudaSetDevice(device));
    // cudaGetDeviceProperties costs ~130 ms per call on this driver; two attributes are all that is needed
    int cc_major = 0, cc_minor = 0, n_sm = 0;
    RR_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
    RR_CUDA(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device));
    RR_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    if (cc_major < 10) { rr_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, cc_major, cc_minor); return RR_E_NODEV; }
    pk->device = device; pk->n_sm = n_sm;
    pk->R = R; pk->N = N; pk->codes = codes;
    pk->row_lo = row_lo; pk->row_hi = row_hi;
    pk->W32 = ((R + 127) / 128) * 4;
    if (pk->W32 == 0) pk->W32 = 4;
    RR_CUDA(cudaStreamCreateWithFlags(&pk->st, cudaStreamNonBlocking));
    rr_alloc_stream(pk->st);
    if (cudaEventCreate(&pk->pe0) != cudaSuccess || cudaEventCreate(&pk->pe1) != cudaSuccess || cudaEventCreate(&pk->pe2) != cudaSuccess) {
        cudaGetLastError(); rr_set_error("cudaEventCreate failed"); return RR_E_CUDA;
    }
    const int n = row_hi - row_lo;
    const size_t ncell = (size_t)n * N;
    int rc;
    if ((rc = dev_alloc(&pk->d_cells, ncell))) return rc;
    RR_TRACE("pack: cells allocated");
    RR_CUDA(cudaEventRecord(pk->pe0, pk->st));
    if (ncell) {
        if (cells && msa->pinned) RR_CUDA(cudaMemcpyAsync(pk->d_cells, cells + (size_t)row_lo * N, ncell, cudaMemcpyHostToDevice, pk->st));
        else if ((rc = staged_upload(msa, row_lo, row_hi, pk->d_cells, device, pk->st))) return rc;
    }
    RR_CUDA(cudaEventRecord(pk->pe1, pk->st));
    RR_TRACE("pack: h2d issued");
    int32_t *d_span = nullptr;
    dev_scope span_scope;                                                // d_span goes back to the pool on every exit path
    if ((rc = span_scope.alloc(&d_span, (size_t)3 * std::max(n, 1)))) return rc;
    RR_CUDA(rr_launch_row_spans(pk->d_cells, n, N, codes, d_span, d_span + n, d_span + 2 * (size_t)n, pk->st));
    pk->slice_spans.assign((size_t)3 * std::max(n, 1), 0);
    RR_CUDA(cudaMemcpyAsync(pk->slice_spans.data(), d_span, sizeof(int32_t) * 3 * (size_t)n, cudaMemcpyDeviceToHost, pk->st));
    RR_CUDA(cudaStreamSynchronize(pk->st));
    RR_TRACE("pack: spans back");
    pk->phase = 1;
    return RR_OK;
}

// spans: start[R], end[R], covered cells[R] of ALL rows of the MSA (in row order)
static int pack_phase_b(rr_packed *pk, const int32_t *sst, const int32_t *sen, const int32_t *scn)
{
    if (pk->phase != 1) { rr_set_error("rr_pack_set_spans: call rr_pack_rows first (once)"); return RR_E_ARG; }
    RR_CUDA(cudaSetDevice(pk->device));
    rr_alloc_stream(pk->st);
    const int R = pk->R, N = pk->N;
    std::vector<int32_t> perm(R);
    std::iota(perm.begin(), perm.end(), 0);
    // Row order = (length class, span start, span end).  The classes cut the rows sorted by span length at fixed
    // fractions, rounded down to whole 256-row K blocks (one class below 1024 rows): see rr_plan.cpp for why rows of
    // different lengths are kept apart.  Any order gives the same counts; this one gives tight K ranges.
    std::vector<uint8_t> cls(std::max(R, 1), 0);
    pk->n_classes = rr_length_classes(R, pk->class_start);
    {
        std::vector<int32_t> by_len(perm);
        std::stable_sort(by_len.begin(), by_len.end(), [&](int32_t a, int32_t b) {
            const int la = sen[a] - sst[a], lb = sen[b] - sst[b];
            if (la != lb) return la < lb;
            return sst[a] < sst[b];
        });
        for (int c = 0; c < pk->n_classes; c++)
            for (int k = pk->class_start[c]; k < pk->class_start[c + 1]; k++) cls[by_len[k]] = (uint8_t)c;
    }
    std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) {
        if (cls[a] != cls[b]) return cls[a] < cls[b];
        if (sst[a] != sst[b]) return sst[a] < sst[b];
        return sen[a] < sen[b];
    });
    pk->h_start.resize(R); pk->h_end.resize(R);
    pk->contiguous = true;
    for (int k = 0; k < R; k++) {
        const int r = perm[k];
        pk->h_start[k] = sst[r]; pk->h_end[k] = sen[r];
        if (scn[r] > 0 && scn[r] != sen[r] - sst[r] + 1) pk->contiguous = false;
    }
    pk->h_perm = perm;
    int rc;
    if ((rc = dev_alloc(&pk->d_perm, (size_t)R))) return rc;
    if (R) RR_CUDA(cudaMemcpyAsync(pk->d_perm, perm.data(), sizeof(int32_t) * R, cudaMemcpyHostToDevice, pk->st));
    // bitsets: one allocation, groups then coverage, so that a merge over slices is one reduction
    const size_t G = (size_t)5 * N;
    if ((rc = dev_alloc(&pk->d_bits, (G + (size_t)N) * pk->W32))) return rc;
    pk->d_covbits = pk->d_bits + G * pk->W32;
    RR_CUDA(rr_launch_pack_bits(pk->d_cells, pk->d_perm, R, N, pk->codes, pk->d_bits, pk->d_covbits, pk->W32, pk->row_lo, pk->row_hi, pk->st));
    RR_CUDA(cudaEventRecord(pk->pe2, pk->st));
    RR_CUDA(cudaStreamSynchronize(pk->st));                              // perm is a local; the caller may merge the bitsets next
    // d_cells is not needed any more (everything downstream works on the bitsets) but stays allocated until the handle is
    // freed: returning 1.8 GB to the stream-ordered pool here and taking 4.6 GB for the operand a moment later made the pool
    // re-map physical memory on every pack when a second handle was alive (measured: +70..1900 ms per rr_pack + rr_scan)
    pk->phase = 2;
    return RR_OK;
}

static int pack_phase_c(rr_packed *pk)
{
    if (pk->phase != 2) { rr_set_error("rr_pack_finish: call rr_pack_set_spans first (once)"); return RR_E_ARG; }
    RR_CUDA(cudaSetDevice(pk->device));
    rr_alloc_stream(pk->st);
    const int R = pk->R, N = pk->N;
    const size_t G = (size_t)5 * N;
    int rc;
    cudaEvent_t e3 = nullptr;
    event_scope ev;
    if (ev.create(&e3)) { rr_set_error("cudaEventCreate failed"); return RR_E_CUDA; }
    RR_CUDA(cudaEventRecord(e3, pk->st));
    if ((rc = dev_alloc(&pk->d_gsize, G))) return rc;
    if ((rc = dev_alloc(&pk->d_coverage, (size_t)N))) return rc;
    RR_CUDA(rr_launch_bitset_sizes(pk->d_bits, (int64_t)G, pk->W32, pk->d_gsize, pk->st));
    RR_CUDA(rr_launch_bitset_sizes(pk->d_covbits, (int64_t)N, pk->W32, pk->d_coverage, pk->st));
    pk->h_gsize.resize(G); pk->h_coverage.resize(N);
    if (G) RR_CUDA(cudaMemcpyAsync(pk->h_gsize.data(), pk->d_gsize, sizeof(int32_t) * G, cudaMemcpyDeviceToHost, pk->st));
    if (N) RR_CUDA(cudaMemcpyAsync(pk->h_coverage.data(), pk->d_coverage, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, pk->st));

    // ln(n!) table, n = 0..R+1
    std::vector<double> lnf((size_t)R + 2);
    rr_lnfact_table(lnf.data(), lnf.size());
    if ((rc = dev_alloc(&pk->d_lnfact, lnf.size()))) return rc;
    RR_CUDA(cudaMemcpyAsync(pk->d_lnfact, lnf.data(), sizeof(double) * lnf.size(), cudaMemcpyHostToDevice, pk->st));

    if ((rc = dev_alloc(&pk->d_best, G))) return rc;
    if ((rc = dev_alloc(&pk->d_counters, (size_t)8))) return rc;
    cudaEvent_t e4 = nullptr;
    if (ev.create(&e4)) { rr_set_error("cudaEventCreate failed"); return RR_E_CUDA; }
    RR_CUDA(cudaEventRecord(e4, pk->st));
    RR_CUDA(cudaStreamSynchronize(pk->st));
    RR_TRACE("pack: done");
    float ab = 0.f, c = 0.f;
    RR_CUDA(cudaEventElapsedTime(&pk->h2d_ms, pk->pe0, pk->pe1));
    RR_CUDA(cudaEventElapsedTime(&ab, pk->pe1, pk->pe2));
    RR_CUDA(cudaEventElapsedTime(&c, e3, e4));
    pk->pack_ms = ab + c;
    if (rr_trace_on()) fprintf(stderr, "[rr trace] h2d %.3f ms, pack %.3f ms\n", pk->h2d_ms, pk->pack_ms);
    pk->phase = 3;
    return RR_OK;
}

extern "C" int rr_pack_rows(const rr_msa *msa, int device, int row_lo, int row_hi, rr_packed **out)
{
    if (!msa || !out) { rr_set_error("rr_pack_rows: bad arguments"); return RR_E_ARG; }
    rr_packed *pk = new rr_packed();
    int rc = pack_phase_a(msa, device, row_lo, row_hi, pk);
    if (rc) { rr_packed_free(pk); *out = nullptr; return rc; }
    *out = pk;
    return RR_OK;
}

extern "C" int rr_pack_slice_spans(rr_packed *pk, int32_t *start, int32_t *end, int32_t *count)
{
    if (!pk || pk->phase < 1 || !start || !end || !count) { rr_set_error("rr_pack_slice_spans: bad arguments"); return RR_E_ARG; }
    const size_t n = (size_t)(pk->row_hi - pk->row_lo);
    memcpy(start, pk->slice_spans.data(), sizeof(int32_t) * n);
    memcpy(end, pk->slice_spans.data() + n, sizeof(int32_t) * n);
    memcpy(count, pk->slice_spans.data() + 2 * n, sizeof(int32_t) * n);
    return RR_OK;
}

extern "C" int rr_pack_set_spans(rr_packed *pk, const int32_t *start, const int32_t *end, const int32_t *count)
{
    if (!pk || !start || !end || !count) { rr_set_error("rr_pack_set_spans: bad arguments"); return RR_E_ARG; }
    return pack_phase_b(pk, start, end, count);
}

extern "C" int rr_pack_bits_device(rr_packed *pk, void **d_bits, size_t *bytes)
{
    if (!pk || pk->phase < 2 || !d_bits || !bytes) { rr_set_error("rr_pack_bits_device: bad arguments"); return RR_E_ARG; }
    *d_bits = pk->d_bits;
    *bytes = (size_t)6 * pk->N * pk->W32 * sizeof(uint32_t);
    return RR_OK;
}

extern "C" int rr_pack_finish(rr_packed *pk)
{
    if (!pk) { rr_set_error("rr_pack_finish: bad arguments"); return RR_E_ARG; }
    return pack_phase_c(pk);
}

extern "C" int rr_pack(const rr_msa *msa, int device, rr_packed **out)
{
    if (!msa || !out) { rr_set_error("rr_pack: bad arguments"); return RR_E_ARG; }
    rr_packed *pk = new rr_packed();
    int rc = pack_phase_a(msa, device, 0, msa->rows, pk);
    if (!rc) {
        const size_t n = (size_t)msa->rows;
        const int32_t *sp = pk->slice_spans.data();
        rc = pack_phase_b(pk, sp, sp + n, sp + 2 * n);
    }
    if (!rc) rc = pack_phase_c(pk);
    if (rc) { rr_packed_free(pk); *out = nullptr; return rc; }
    *out = pk;
    return RR_OK;
}

extern "C" int rr_timer_start(rr_packed *pk)
{
    if (!pk) return RR_E_ARG;
    RR_CUDA(cudaSetDevice(pk->device));
    if (!pk->t0) { RR_CUDA(cudaEventCreate(&pk->t0)); RR_CUDA(cudaEventCreate(&pk->t1)); }
    RR_CUDA(cudaEventRecord(pk->t0, pk->st));
    return RR_OK;
}

extern "C" int rr_timer_stop(rr_packed *pk, float *elapsed_ms)
{
    if (!pk || !pk->t0 || !elapsed_ms) return RR_E_ARG;
    RR_CUDA(cudaSetDevice(pk->device));
    RR_CUDA(cudaEventRecord(pk->t1, pk->st));
    RR_CUDA(cudaEventSynchronize(pk->t1));
    RR_CUDA(cudaEventElapsedTime(elapsed_ms, pk->t0, pk->t1));
    return RR_OK;
}

extern "C" int rr_packed_sizes(rr_packed *pk, int32_t *gsize, int32_t *coverage)
{
    if (!pk || pk->phase != 3) { rr_set_error("rr_packed_sizes: the packed MSA is incomplete"); return RR_E_ARG; }
    if (gsize) memcpy(gsize, pk->h_gsize.data(), sizeof(int32_t) * pk->h_gsize.size());
    if (coverage) memcpy(coverage, pk->h_coverage.data(), sizeof(int32_t) * pk->h_coverage.size());
    return RR_OK;
}

extern "C" int rr_pair_counts(rr_packed *pk, int64_t n, const int32_t *gi, const int32_t *gj, int32_t *out)
{
    if (!pk || pk->phase != 3 || n < 0 || (n && (!gi || !gj || !out))) { rr_set_error("rr_pair_counts: bad arguments"); return RR_E_ARG; }
    if (n == 0) return RR_OK;
    RR_CUDA(cudaSetDevice(pk->device));
    rr_alloc_stream(pk->st);
    dev_scope scope;
    int32_t *d_i = nullptr, *d_j = nullptr, *d_o = nullptr;
    int rc;
    if ((rc = scope.alloc(&d_i, (size_t)n)) || (rc = scope.alloc(&d_j, (size_t)n)) || (rc = scope.alloc(&d_o, (size_t)4 * n))) return rc;
    RR_CUDA(cudaMemcpyAsync(d_i, gi, sizeof(int32_t) * n, cudaMemcpyHostToDevice, pk->st));
    RR_CUDA(cudaMemcpyAsync(d_j, gj, sizeof(int32_t) * n, cudaMemcpyHostToDevice, pk->st));
    RR_CUDA(rr_launch_pair_counts(pk->d_bits, pk->d_covbits, pk->W32, n, d_i, d_j, d_o, pk->st));
    RR_CUDA(cudaMemcpyAsync(out, d_o, sizeof(int32_t) * 4 * n, cudaMemcpyDeviceToHost, pk->st));
    RR_CUDA(cudaStreamSynchronize(pk->st));
    return RR_OK;
}

// ---------------------------------------------------------------------------------------
// Cliquer (RepeatResolver.c:1179-1240), first version: device counts, host scores
// ---------------------------------------------------------------------------------------
extern "C" double rr_group_score_host(uint32_t s, uint32_t gr1, uint32_t gr2, uint32_t cov, int32_t sizei, int32_t sizej)
{
    std::vector<double> &t = host_lnfact((size_t)cov + 2);
    return rr_group_significance(t.data(), s, gr1, gr2, cov, sizei, sizej);
}

// scores the candidates with the host libm and applies TheBestUpdater's order (1156-1176): the maxclique-1 largest
// scores above greedy, equal scores in candidate order (candidates are given in ascending group order)
static void clq_select(int query_group, int64_t n, const int32_t *groups, const int32_t *counts, const int32_t *sizes,
                       int size_query, int mincov, int maxclique, double greedy, int threads, int32_t *members,
                       double *scores, int *n_members)
{
    for (int j = 0; j <= maxclique; j++) members[j] = -1;
    for (int j = 0; j < maxclique; j++) scores[j] = 0.0;
    members[0] = query_group;                                            // 1197
    scores[0] = 100.0;                                                   // 1229
    std::vector<double> Z((size_t)n, 0.0);
    int32_t max_cov = 0;
    for (int64_t k = 0; k < n; k++) max_cov = std::max(max_cov, counts[4 * k + 3]);
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)threads, (int64_t)std::thread::hardware_concurrency(), n / 4096 + 1}));
    auto work = [&](int t) {
        host_lnfact((size_t)max_cov + 2);   // per-thread table, grown once
        for (int64_t k = n * t / nt; k < n * (t + 1) / nt; k++) {
            const int32_t *c = counts + 4 * k;
            if (groups[k] == query_group || c[0] <= mincov / 4) continue;     // 1210, 1215
            Z[k] = rr_group_score_host((uint32_t)c[0], (uint32_t)c[1], (uint32_t)c[2], (uint32_t)c[3], sizes[k], size_query);
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    std::vector<int64_t> cand;
    for (int64_t k = 0; k < n; k++)
        if (Z[k] > greedy) cand.push_back(k);                            // 1218
    std::stable_sort(cand.begin(), cand.end(), [&](int64_t a, int64_t b) { return Z[a] > Z[b]; });
    int m = 1;
    for (size_t r = 0; r < cand.size() && m < maxclique; r++, m++) {
        members[m] = groups[cand[r]];
        scores[m] = Z[cand[r]];
    }
    *n_members = m;
}

extern "C" int rr_cliquer_from_counts(int query_group, int64_t n, const int32_t *groups, const int32_t *counts,
                                      const int32_t *sizes, int size_query, int mincov, int maxclique, double greedy,
                                      int32_t *members, double *scores, int *n_members)
{
    if (n < 0 || maxclique < 1 || mincov < 0 || !(greedy >= 0.0) || !members || !scores || !n_members || (n && (!groups || !counts || !sizes))) {
        rr_set_error("rr_cliquer_from_counts: bad arguments");
        return RR_E_ARG;
    }
    for (int64_t k = 1; k < n; k++)
        if (groups[k] <= groups[k - 1]) { rr_set_error("rr_cliquer_from_counts: candidate groups must be ascending"); return RR_E_ARG; }
    clq_select(query_group, n, groups, counts, sizes, size_query, mincov, maxclique, greedy, 16, members, scores, n_members);
    return RR_OK;
}

// one query, every candidate's counts from rr_pair_counts, every score on the host: the plain second implementation
// the tests hold rr_cliquer_batch against
// ---------------------------------------------------------------------------------------
// CliqueGroup / CliqueCoverage for a batch of cliques (RepeatResolver.c:976-1008, 1064-1096; called 1662-1664)
// ---------------------------------------------------------------------------------------
extern "C" int rr_clique_groups(rr_packed *pk, int64_t n_cliques, const int32_t *members, int stride, const int32_t *n_members,
                                const int32_t *cutoffs, uint64_t *groups, uint64_t *coverage)
{
    if (!pk || pk->phase != 3 || n_cliques < 0 || stride < 1 || (n_cliques && (!members || !n_members || !cutoffs))) {
        rr_set_error("rr_clique_groups: bad arguments");
        return RR_E_ARG;
    }
    if ((int)pk->h_perm.size() != pk->R) { rr_set_error("rr_clique_groups: the packed MSA carries no row order"); return RR_E_ARG; }
    for (int64_t q = 0; q < n_cliques; q++) {
        // the reference sizes a clique by its first negative entry among 100 (986-993)
        if (n_members[q] < 0 || n_members[q] > stride || n_members[q] > 100) { rr_set_error("rr_clique_groups: clique %lld has %d members (0..min(stride, 100))", (long long)q, n_members[q]); return RR_E_ARG; }
        for (int m = 0; m < n_members[q]; m++)
            if (members[q * stride + m] < 0 || members[q * stride + m] >= 5 * pk->N) { rr_set_error("rr_clique_groups: group %d out of range", members[q * stride + m]); return RR_E_ARG; }
    }
    const int R = pk->R;
    const size_t sc = (size_t)R / 64 + 1;                                // words of a group (RepeatResolver.c:59)
    if (n_cliques == 0 || (!groups && !coverage)) return RR_OK;
    RR_CUDA(cudaSetDevice(pk->device));
    rr_alloc_stream(pk->st);
    dev_scope scope;
    int rc;
    int32_t *d_members = nullptr, *d_nm = nullptr, *d_cut = nullptr, *d_rank = nullptr;
    uint32_t *d_tmp = nullptr, *d_out = nullptr;
    // cliques in slices that keep the device buffers small
    const int64_t slice = std::max<int64_t>(1, std::min<int64_t>(n_cliques, ((int64_t)256 << 20) / (int64_t)(8 * sc + 4 * pk->W32)));
    if ((rc = scope.alloc(&d_members, (size_t)slice * stride)) || (rc = scope.alloc(&d_nm, (size_t)slice)) || (rc = scope.alloc(&d_cut, (size_t)slice)) ||
        (rc = scope.alloc(&d_rank, (size_t)std::max(R, 1))) || (rc = scope.alloc(&d_tmp, (size_t)slice * pk->W32)) || (rc = scope.alloc(&d_out, (size_t)slice * 2 * sc)))
        return rc;
    std::vector<int32_t> rank_of_row(std::max(R, 1), 0);
    for (int r = 0; r < R; r++) rank_of_row[pk->h_perm[r]] = r;
    RR_CUDA(cudaMemcpyAsync(d_rank, rank_of_row.data(), sizeof(int32_t) * std::max(R, 1), cudaMemcpyHostToDevice, pk->st));
    for (int64_t q0 = 0; q0 < n_cliques; q0 += slice) {
        const int64_t n = std::min(slice, n_cliques - q0);
        RR_CUDA(cudaMemcpyAsync(d_members, members + q0 * stride, sizeof(int32_t) * n * stride, cudaMemcpyHostToDevice, pk->st));
        RR_CUDA(cudaMemcpyAsync(d_nm, n_members + q0, sizeof(int32_t) * n, cudaMemcpyHostToDevice, pk->st));
        RR_CUDA(cudaMemcpyAsync(d_cut, cutoffs + q0, sizeof(int32_t) * n, cudaMemcpyHostToDevice, pk->st));
        for (int which = 0; which < 2; which++) {
            uint64_t *dst = which ? coverage : groups;
            if (!dst) continue;
            RR_CUDA(rr_launch_clique_members(pk->d_bits, pk->d_covbits, pk->W32, n, d_members, stride, d_nm, d_cut, which, d_rank, R,
                                             (int)(2 * sc), d_tmp, d_out, pk->st));
            RR_CUDA(cudaMemcpyAsync(dst + q0 * sc, d_out, sizeof(uint64_t) * n * sc, cudaMemcpyDeviceToHost, pk->st));
            RR_CUDA(cudaStreamSynchronize(pk->st));
        }
    }
    return RR_OK;
}

extern "C" int rr_cliquer(rr_packed *pk, int query_group, int anfang, int ende, int mincov, int maxclique, double greedy,
                          int32_t *members, double *scores, int *n_members)
{
    // greedy >= 0: TheBestUpdater only ever replaces entries of Best_Corrs that start at 0.0 (1156-1176), so no score <= 0 can
    // enter a clique; a negative greedy would let the candidates skipped by 1210/1215 (scored 0 here) in
    if (!pk || maxclique < 1 || mincov < 0 || !(greedy >= 0.0) || !members || !scores || !n_members) { rr_set_error("rr_cliquer: bad arguments"); return RR_E_ARG; }
    if (query_group < 0 || query_group >= 5 * pk->N) { rr_set_error("rr_cliquer: group %d out of range", query_group); return RR_E_ARG; }
    anfang = std::max(anfang, 0);
    ende = std::min(ende, pk->N);
    const int64_t n = ende > anfang ? (int64_t)5 * (ende - anfang) : 0;
    std::vector<int32_t> gi((size_t)n), gj((size_t)n, query_group), cnt((size_t)4 * n), sz((size_t)n);
    for (int64_t k = 0; k < n; k++) {
        gi[k] = (int32_t)(5 * (int64_t)anfang + k);                      // Group1 = the candidate, Group2 = the query (1217)
        sz[k] = pk->h_gsize[gi[k]];
    }
    int rc = rr_pair_counts(pk, n, gi.data(), gj.data(), cnt.data());
    if (rc) return rc;
    clq_select(query_group, n, gi.data(), cnt.data(), sz.data(), pk->h_gsize[query_group], mincov, maxclique, greedy, 16,
               members, scores, n_members);
    return RR_OK;
}

// Cliquer for a batch of query groups (the calls of Group_Refinement, 1647-1649): counts, the 1215 filter, a rigorous
// score bound and the exact score on the device (rr_cliquer.cu); the host keeps, per query, the hits that can decide
// the top maxclique-1, re-evaluates them with its own libm and orders them like TheBestUpdater.  A device score
// differs from the host's by a few ulp (exp, log10), so "can decide" = within 1e-9 relative of the weakest of the
// top maxclique-1 device scores, or anywhere near the saturation switch at 98 (486), where a last-bit difference
// selects the other formula.
// The host half of rr_cliquer_batch for the n queries of one launch: `hits` holds, in no particular order, the listed
// pairs with their counts and device scores; slot = index into queries.  Per query the hits that can decide the top
// maxclique-1 are re-evaluated and ordered by clq_select.
static void clq_finalize(const int32_t *queries, int n, std::vector<rr_clq_rec> &hits, const int32_t *gsize, int mincov,
                         int maxclique, double greedy, int32_t *members, double *scores, int32_t *n_members, long long *evals_out)
{
    const int K = maxclique - 1;
    // bucket the hits by query slot (one linear pass; the device appends them in no particular order)
    std::vector<int64_t> first((size_t)n + 1, 0);
    for (const rr_clq_rec &h : hits) first[(size_t)h.slot + 1]++;
    for (int i = 0; i < n; i++) first[(size_t)i + 1] += first[i];
    std::vector<rr_clq_rec> by_slot(hits.size());
    {
        std::vector<int64_t> fill(first.begin(), first.end() - 1);
        for (const rr_clq_rec &h : hits) by_slot[(size_t)fill[h.slot]++] = h;
    }
    std::atomic<int> next{0};
    std::atomic<long long> evals{0};
    auto work = [&]() {
        std::vector<double> zs;
        std::vector<rr_clq_rec> keep;
        std::vector<int32_t> g, c, sz;
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) return;
            const rr_clq_rec *h = by_slot.data() + first[i];
            const int64_t nh = first[(size_t)i + 1] - first[i];
            if (nh == 0 || K < 1) continue;
            double cut = -HUGE_VAL;
            if (nh > K) {
                zs.resize((size_t)nh);
                for (int64_t k = 0; k < nh; k++) zs[k] = h[k].z;
                std::nth_element(zs.begin(), zs.begin() + (K - 1), zs.end(), std::greater<double>());
                cut = zs[K - 1] - 1e-9 * std::max(1.0, std::fabs(zs[K - 1]));
            }
            keep.clear();
            for (int64_t k = 0; k < nh; k++)
                if (!(h[k].z < cut && h[k].z < 97.89)) keep.push_back(h[k]);
            // candidates in ascending group order: equal scores keep the earlier group (TheBestUpdater, 1156-1176)
            std::sort(keep.begin(), keep.end(), [](const rr_clq_rec &a, const rr_clq_rec &b) { return a.group < b.group; });
            g.clear(); c.clear(); sz.clear();
            for (const rr_clq_rec &r : keep) {
                g.push_back(r.group);
                c.push_back(r.s); c.push_back(r.gr1); c.push_back(r.gr2); c.push_back(r.cov);
                sz.push_back(gsize[r.group]);
            }
            int m = 0;
            clq_select(queries[i], (int64_t)g.size(), g.data(), c.data(), sz.data(), gsize[queries[i]], mincov, maxclique, greedy, 1,
                       members + i * ((int64_t)maxclique + 1), scores + i * (int64_t)maxclique, &m);
            n_members[i] = m;
            evals += (long long)g.size();
        }
    };
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)16, (int64_t)std::thread::hardware_concurrency(), (int64_t)n / 8 + 1}));
    if (nt == 1) work();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    if (evals_out) *evals_out = evals.load();
}

static void clq_init_outputs(int64_t nq, const int32_t *queries, int maxclique, int32_t *members, double *scores, int32_t *n_members)
{
    for (int64_t i = 0; i < nq; i++) {                                   // the clique of a query without partners
        int32_t *m = members + i * ((int64_t)maxclique + 1);
        double *z = scores + i * (int64_t)maxclique;
        for (int j = 0; j <= maxclique; j++) m[j] = -1;
        for (int j = 0; j < maxclique; j++) z[j] = 0.0;
        m[0] = queries[i];
        z[0] = 100.0;
        n_members[i] = 1;
    }
}

// test hook: clq_finalize on hits given by the caller (layout of rr_clq_rec: six int32 {slot, group, s, gr1, gr2, cov}
// and the device score as a double)
extern "C" int rr_cliquer_from_hits(int64_t nq, const int32_t *queries, int64_t n_hits, const void *hit_records,
                                    const int32_t *gsize, int64_t n_groups, int mincov, int maxclique, double greedy,
                                    int32_t *members, double *scores, int32_t *n_members)
{
    if (nq < 0 || nq > 0x7fffffff || n_hits < 0 || maxclique < 1 || mincov < 0 || !(greedy >= 0.0) || !gsize || (n_hits && !hit_records) ||
        (nq && (!queries || !members || !scores || !n_members))) {
        rr_set_error("rr_cliquer_from_hits: bad arguments");
        return RR_E_ARG;
    }
    const rr_clq_rec *h = static_cast<const rr_clq_rec *>(hit_records);
    for (int64_t i = 0; i < nq; i++)
        if (queries[i] < 0 || queries[i] >= n_groups) { rr_set_error("rr_cliquer_from_hits: query out of range"); return RR_E_ARG; }
    for (int64_t k = 0; k < n_hits; k++)
        if (h[k].slot < 0 || h[k].slot >= nq || h[k].group < 0 || h[k].group >= n_groups) {
            rr_set_error("rr_cliquer_from_hits: hit %lld out of range", (long long)k);
            return RR_E_ARG;
        }
    clq_init_outputs(nq, queries, maxclique, members, scores, n_members);
    std::vector<rr_clq_rec> hits(h, h + n_hits);
    clq_finalize(queries, (int)nq, hits, gsize, mincov, maxclique, greedy, members, scores, n_members, nullptr);
    return RR_OK;
}

// capacity of the candidate / hit lists of rr_cliquer_batch (records of 32 bytes); tests lower it to force the retry path
static std::atomic<unsigned long long> g_cliquer_cap{1ull << 24};
// capacity (candidates) of the deferred-evaluation list of the tcgen05 scan; 0 = sized from the plan.  Tests lower it so
// that the list overflows and the kernel's in-place evaluation takes over mid-scan.
static unsigned long long g_deferred_cap = 0;
extern "C" void rr_debug_set_deferred_cap(unsigned long long cap) { g_deferred_cap = cap; }

extern "C" void rr_debug_set_cliquer_cap(unsigned long long cap) { g_cliquer_cap = cap ? cap : (1ull << 24); }

static void clq_release(int32_t *q, unsigned long long *c, rr_clq_rec *a, rr_clq_rec *b, cudaEvent_t e0, cudaEvent_t e1)
{
    rr_dev_free(q); rr_dev_free(c); rr_dev_free(a); rr_dev_free(b);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
}

extern "C" int rr_cliquer_batch(rr_packed *pk, int64_t nq, const int32_t *queries, int anfang, int ende, int mincov,
                                int maxclique, double greedy, int32_t *members, double *scores, int32_t *n_members,
                                rr_cliquer_stats *stats)
{
    if (stats) memset(stats, 0, sizeof(*stats));
    if (!pk || pk->phase != 3 || nq < 0 || nq > 0x7fffffff || maxclique < 1 || mincov < 0 || !(greedy >= 0.0) ||
        (nq && (!queries || !members || !scores || !n_members))) {
        rr_set_error("rr_cliquer_batch: bad arguments");
        return RR_E_ARG;
    }
    for (int64_t i = 0; i < nq; i++)
        if (queries[i] < 0 || queries[i] >= 5 * pk->N) { rr_set_error("rr_cliquer_batch: group %d out of range", queries[i]); return RR_E_ARG; }
    clq_init_outputs(nq, queries, maxclique, members, scores, n_members);
    anfang = std::max(anfang, 0);
    ende = std::min(ende, pk->N);
    if (nq == 0 || ende <= anfang || maxclique == 1) return RR_OK;
    const unsigned long long n_cand_groups = 5ull * (unsigned long long)(ende - anfang);
    if ((ende - anfang + RR_CLQ_SLAB - 1) / RR_CLQ_SLAB > 65535 || rr_cliquer_smem_bytes(pk->W32) > 227u * 1024u) {
        rr_set_error("rr_cliquer_batch: MSA too large for the kernel's tiling (%d columns, %d rows)", ende - anfang, pk->R);
        return RR_E_ARG;
    }
    RR_CUDA(cudaSetDevice(pk->device));
    rr_alloc_stream(pk->st);

    unsigned long long cap = g_cliquer_cap;                              // entries of 32 bytes per list
    int64_t group_len = std::min<int64_t>(nq, 4096);
    cap = std::min(cap, (unsigned long long)group_len * n_cand_groups);

    int32_t *d_queries = nullptr;
    unsigned long long *d_counters = nullptr;
    rr_clq_rec *d_cand = nullptr, *d_hits = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int rc;
    if ((rc = dev_alloc(&d_queries, (size_t)nq)) || (rc = dev_alloc(&d_counters, 2)) || (rc = dev_alloc(&d_cand, (size_t)cap)) ||
        (rc = dev_alloc(&d_hits, (size_t)cap))) {
        clq_release(d_queries, d_counters, d_cand, d_hits, ev0, ev1);
        return rc;
    }
#define CLQ_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            rr_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
            clq_release(d_queries, d_counters, d_cand, d_hits, ev0, ev1);                          \
            return RR_E_CUDA;                                                                      \
        }                                                                                          \
    } while (0)
    CLQ_CUDA(cudaEventCreate(&ev0));
    CLQ_CUDA(cudaEventCreate(&ev1));
    CLQ_CUDA(cudaMemcpyAsync(d_queries, queries, sizeof(int32_t) * (size_t)nq, cudaMemcpyHostToDevice, pk->st));

    double thr = greedy - 1e-9 * std::max(1.0, std::fabs(greedy));
    thr = std::min(thr, 97.89);                                          // both sides of the switch at 98 reach the host
    std::vector<rr_clq_rec> hits;
    int64_t q0 = 0;
    while (q0 < nq) {
        const int n = (int)std::min<int64_t>(group_len, nq - q0);
        unsigned long long cnt[2] = {0, 0};
        float ms = 0.f;
        CLQ_CUDA(cudaMemsetAsync(d_counters, 0, 2 * sizeof(unsigned long long), pk->st));
        CLQ_CUDA(cudaEventRecord(ev0, pk->st));
        CLQ_CUDA(rr_launch_cliquer(pk->d_bits, pk->d_covbits, pk->d_gsize, pk->d_lnfact, pk->W32, d_queries + q0, n, anfang, ende,
                                   mincov / 4, greedy, thr, d_cand, d_hits, cap, d_counters, pk->n_sm, pk->st));
        CLQ_CUDA(cudaEventRecord(ev1, pk->st));
        CLQ_CUDA(cudaMemcpyAsync(cnt, d_counters, sizeof(cnt), cudaMemcpyDeviceToHost, pk->st));
        CLQ_CUDA(cudaStreamSynchronize(pk->st));
        CLQ_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        if (stats) { stats->kernel_ms += ms; stats->launches += 2; }
        if (cnt[0] > cap) {                                              // candidate list overflowed: fewer queries per launch
            if (stats) stats->retries++;
            if (n > 1) { group_len = std::max(1, n / 2); continue; }
            cap = n_cand_groups;                                         // one query lists at most every candidate group
            rr_dev_free(d_cand); rr_dev_free(d_hits);
            d_cand = d_hits = nullptr;
            if ((rc = dev_alloc(&d_cand, (size_t)cap)) || (rc = dev_alloc(&d_hits, (size_t)cap))) {
                clq_release(d_queries, d_counters, d_cand, d_hits, ev0, ev1);
                return rc;
            }
            continue;
        }
        hits.resize((size_t)cnt[1]);
        if (cnt[1]) {
            CLQ_CUDA(cudaMemcpyAsync(hits.data(), d_hits, sizeof(rr_clq_rec) * (size_t)cnt[1], cudaMemcpyDeviceToHost, pk->st));
            CLQ_CUDA(cudaStreamSynchronize(pk->st));
        }
        if (stats) {
            stats->pairs += (int64_t)n * (int64_t)n_cand_groups;
            for (int i = 0; i < n; i++)                                  // 1210: a query group is not its own candidate
                if (queries[q0 + i] / 5 >= anfang && queries[q0 + i] / 5 < ende) stats->pairs--;
            stats->candidates += (int64_t)cnt[0];
            stats->hits += (int64_t)cnt[1];
        }
        long long evals = 0;
        clq_finalize(queries + q0, n, hits, pk->h_gsize.data(), mincov, maxclique, greedy, members + q0 * ((int64_t)maxclique + 1),
                     scores + q0 * (int64_t)maxclique, n_members + q0, &evals);
        if (stats) stats->host_evals += evals;
        q0 += n;
    }
#undef CLQ_CUDA
    clq_release(d_queries, d_counters, d_cand, d_hits, ev0, ev1);
    return RR_OK;
}

// ---------------------------------------------------------------------------------------
// Group_Refinement (RepeatResolver.c:1634-1693; Parallel_Group_Refinement 1770-1821 does the same group by group): Cliquer
// for every group above the cutoff in one device batch, Sizes (1650), Dropoff_Cutoff(i, 0) on the device's member counts -
// BestCutoff and KorrMaxCutoff run before it in the reference, but their results are overwritten at 1662 and they have no
// side effect - then CliqueGroup and CliqueCoverage at that cutoff; MaxCorrs zeroed where Sizes <= 5 (1685).
// ---------------------------------------------------------------------------------------
// Dropoff_Cutoff 1488-1509 on sizes[k] = reads contained in more than k of the clique's first `size` members
extern "C" int rr_dropoff_cutoff_host(const uint32_t *sizes, int size, int signumber, int c, double *drop_off)
{
    const int c0 = std::max(1, c);
    int drop_c = c0;
    double min_drop = 1000000.0;
    for (int i = c0; i < size - 1; i++) {
        const double si = (double)sizes[i];
        const double den = std::min((double)signumber - si, si);
        if (den > 0) {
            const double drop = ((double)sizes[i - 1] - (double)sizes[i + 1]) / den;
            if (drop < min_drop) { min_drop = drop; drop_c = i; }       // the first minimum wins
        }
    }
    if (drop_off) *drop_off = min_drop;
    return drop_c;
}

extern "C" int rr_group_refinement(rr_packed *pk, double *maxcorrs, double cutoff, int anfang, int ende, int mincov, int maxclique,
                                   double greedy, int64_t capacity, int32_t *query_groups, int32_t *cliques, int32_t *sizes,
                                   int32_t *cutoffs, double *drop_off, uint64_t *c_groups, uint64_t *c_coverage,
                                   int64_t *n_queries, rr_cliquer_stats *stats)
{
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_queries) *n_queries = 0;
    // Dropoff_Cutoff keeps one group per member in an array of 100 (1463)
    if (!pk || pk->phase != 3 || !n_queries || capacity < 0 || maxclique < 1 || maxclique > 100 || mincov < 0 || !(greedy >= 0.0) ||
        (pk->N > 0 && !maxcorrs) || (capacity && (!query_groups || !cliques || !sizes || !cutoffs || !drop_off))) {
        rr_set_error("rr_group_refinement: bad arguments");
        return RR_E_ARG;
    }
    if ((int)pk->h_perm.size() != pk->R) { rr_set_error("rr_group_refinement: the packed MSA carries no row order"); return RR_E_ARG; }
    const int64_t G = (int64_t)5 * pk->N;
    int64_t nq = 0;
    for (int64_t i = 0; i < G; i++)
        if (maxcorrs[i] > cutoff) {                                      // 1647
            if (nq < capacity) query_groups[nq] = (int32_t)i;
            nq++;
        }
    *n_queries = nq;
    if (nq > capacity) {
        rr_set_error("rr_group_refinement: %lld groups above the cutoff, room for %lld", (long long)nq, (long long)capacity);
        return RR_E_ARG;
    }
    const int stride = maxclique + 1;
    const size_t sc = (size_t)pk->R / 64 + 1;
    if (c_groups) memset(c_groups, 0, sizeof(uint64_t) * (size_t)nq * sc);
    if (c_coverage) memset(c_coverage, 0, sizeof(uint64_t) * (size_t)nq * sc);
    if (nq == 0) return RR_OK;
    int rc;
    {
        std::vector<double> scores((size_t)nq * maxclique);
        std::vector<int32_t> nmem((size_t)nq);
        if ((rc = rr_cliquer_batch(pk, nq, query_groups, anfang, ende, mincov, maxclique, greedy, cliques, scores.data(), nmem.data(), stats)))
            return rc;
    }
    std::vector<int64_t> refined;                                        // slots with Sizes > 5 (1652)
    for (int64_t q = 0; q < nq; q++) {
        const int32_t *m = cliques + q * stride;
        int n = 0;
        while (n < maxclique && m[n] > 0) n++;                           // 1650: group 0 ends the count like the -1 does
        sizes[q] = n;
        cutoffs[q] = 0;
        drop_off[q] = 1000.0;                                            // 1645
        if (n > 5) refined.push_back(q);
        else maxcorrs[query_groups[q]] = 0.0;                            // 1685
    }
    const int64_t nr = (int64_t)refined.size();
    if (nr == 0) return RR_OK;
    // member lists of the refined cliques: Dropoff_Cutoff sees the first Sizes members, CliqueGroup / CliqueCoverage every
    // member up to the first negative entry (986-993)
    std::vector<int32_t> mem((size_t)nr * stride), n_drop((size_t)nr), n_all((size_t)nr), cut((size_t)nr);
    for (int64_t k = 0; k < nr; k++) {
        const int32_t *m = cliques + refined[k] * stride;
        memcpy(&mem[(size_t)k * stride], m, sizeof(int32_t) * stride);
        int n = 0;
        while (n < stride && m[n] >= 0) n++;
        n_drop[k] = sizes[refined[k]];
        n_all[k] = n;
    }
    RR_CUDA(cudaSetDevice(pk->device));
    rr_alloc_stream(pk->st);
    {
        dev_scope scope;
        int32_t *d_mem = nullptr, *d_n = nullptr;
        uint32_t *d_sizes = nullptr;
        const int64_t slice = std::min<int64_t>(nr, (int64_t)1 << 20);
        if ((rc = scope.alloc(&d_mem, (size_t)slice * stride)) || (rc = scope.alloc(&d_n, (size_t)slice)) ||
            (rc = scope.alloc(&d_sizes, (size_t)slice * stride)))
            return rc;
        std::vector<uint32_t> h_sizes((size_t)slice * stride);
        for (int64_t k0 = 0; k0 < nr; k0 += slice) {
            const int64_t n = std::min(slice, nr - k0);
            RR_CUDA(cudaMemcpyAsync(d_mem, &mem[(size_t)k0 * stride], sizeof(int32_t) * n * stride, cudaMemcpyHostToDevice, pk->st));
            RR_CUDA(cudaMemcpyAsync(d_n, &n_drop[k0], sizeof(int32_t) * n, cudaMemcpyHostToDevice, pk->st));
            RR_CUDA(rr_launch_clique_sizes(pk->d_bits, pk->W32, n, d_mem, stride, d_n, d_sizes, pk->st));
            RR_CUDA(cudaMemcpyAsync(h_sizes.data(), d_sizes, sizeof(uint32_t) * n * stride, cudaMemcpyDeviceToHost, pk->st));
            RR_CUDA(cudaStreamSynchronize(pk->st));
            if (stats) stats->launches += (int)((n + 65534) / 65535);
            for (int64_t k = 0; k < n; k++) {
                const int64_t q = refined[k0 + k];
                cut[k0 + k] = cutoffs[q] = rr_dropoff_cutoff_host(&h_sizes[(size_t)k * stride], sizes[q], pk->R, 0, &drop_off[q]);  // 1662
            }
        }
    }
    if (!c_groups && !c_coverage) return RR_OK;
    std::vector<uint64_t> g(c_groups ? (size_t)nr * sc : 0), v(c_coverage ? (size_t)nr * sc : 0);
    if ((rc = rr_clique_groups(pk, nr, mem.data(), stride, n_all.data(), cut.data(), c_groups ? g.data() : nullptr,
                               c_coverage ? v.data() : nullptr)))                                                        // 1663, 1665
        return rc;
    for (int64_t k = 0; k < nr; k++) {
        if (c_groups) memcpy(c_groups + (size_t)refined[k] * sc, &g[(size_t)k * sc], sizeof(uint64_t) * sc);
        if (c_coverage) memcpy(c_coverage + (size_t)refined[k] * sc, &v[(size_t)k * sc], sizeof(uint64_t) * sc);
    }
    return RR_OK;
}

// ---------------------------------------------------------------------------------------
// Relative_Vars (RepeatResolver.c:2424-2493): the masked Gram matrix X^T diag(u) X over the selected groups.  Either the
// rows of the part are packed as an MSA of their own (rr_relative_vars: the triple intersections |Gi & Gj & U| become plain
// pair intersections and |Gi & U| its group sizes), or the part is a mask over the packed copy of the whole MSA
// (rr_relative_vars_packed).  Selection on the host; counts, the two-sided score bound, the exact score and the marks in
// the tiled kernel of rr_relvars.cu; pairs within 1e-9 of the cutoff are decided with the host libm.
// ---------------------------------------------------------------------------------------
extern "C" double rr_relative_score_host(uint32_t s, uint32_t gr1, uint32_t gr2, uint32_t cov)
{
    std::vector<double> &t = host_lnfact((size_t)cov + 2);
    return rr_relative_significance(t.data(), s, gr1, gr2, cov);
}

// 2430-2452: groups with MaxCorrs > cutoff of which at least mingroup reads lie in the part
static void relvars_select(int64_t n_groups, const double *maxcorrs, const int32_t *gsize_u, double cutoff, int mingroup,
                           std::vector<int32_t> &sel)
{
    sel.clear();
    for (int64_t g = 0; g < n_groups; g++)
        if (maxcorrs[g] > cutoff && gsize_u[g] >= mingroup) sel.push_back((int32_t)g);
}

// first selected index whose group id is at least sel[a] + 100 (2461)
static size_t relvars_first_partner(const std::vector<int32_t> &sel, size_t a)
{
    return (size_t)(std::lower_bound(sel.begin() + a, sel.end(), sel[a] + 100) - sel.begin());
}

// 2461-2475 for the listed pairs: pa/pb index into sel, S = |G_pa & G_pb & U|
static void relvars_mark(const std::vector<int32_t> &sel, int64_t n, const int32_t *pa, const int32_t *pb, const int32_t *S,
                         int64_t S_stride, const int32_t *gsize_u, int cov_u, double cutoff, std::vector<uint8_t> &mark)
{
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)16, (int64_t)std::thread::hardware_concurrency(), n / 2048 + 1}));
    std::vector<std::vector<uint8_t>> local(nt, std::vector<uint8_t>(nt > 1 ? sel.size() : 0, 0));
    auto work = [&](int t) {
        std::vector<uint8_t> &m = nt > 1 ? local[t] : mark;
        host_lnfact((size_t)cov_u + 2);
        for (int64_t k = n * t / nt; k < n * (t + 1) / nt; k++) {
            const int a = pa[k], b = pb[k];
            if (m[a] && m[b]) continue;
            // Relative_Group_Significance(Groups[j], Groups[i], U_Group): Group1 = the later group (2465)
            const double Z = rr_relative_score_host((uint32_t)S[k * S_stride], (uint32_t)gsize_u[sel[b]], (uint32_t)gsize_u[sel[a]],
                                                    (uint32_t)cov_u);
            if (Z > cutoff) m[a] = m[b] = 1;
        }
    };
    if (nt == 1) { work(0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto &t : th) t.join();
    for (int t = 0; t < nt; t++)
        for (size_t k = 0; k < sel.size(); k++) mark[k] |= local[t][k];
}

// test hook: the host half on given counts.  Call once with S = NULL for the selection (sel_out[n_sel], ascending group
// ids), then with S[n_sel][n_sel] = |G_sel[a] & G_sel[b] & U| for the result.
extern "C" int rr_relative_vars_from_counts(int64_t n_groups, const double *maxcorrs, const int32_t *gsize_u, int cov_u,
                                            double cutoff, int mingroup, const int32_t *S, int32_t *sel_out, int *n_sel,
                                            int32_t *vars, int *n_vars)
{
    if (n_groups < 0 || n_groups > 0x7fffffff || cov_u < 0 || mingroup < 1 || !(cutoff >= 0.0) || !n_sel ||
        (n_groups && (!maxcorrs || !gsize_u))) {
        rr_set_error("rr_relative_vars_from_counts: bad arguments");
        return RR_E_ARG;
    }
    std::vector<int32_t> sel;
    relvars_select(n_groups, maxcorrs, gsize_u, cutoff, mingroup, sel);
    *n_sel = (int)sel.size();
    if (sel_out) std::copy(sel.begin(), sel.end(), sel_out);
    if (!S) return RR_OK;
    if (!vars || !n_vars) { rr_set_error("rr_relative_vars_from_counts: bad arguments"); return RR_E_ARG; }
    std::vector<int32_t> pa, pb, sv;
    for (size_t a = 0; a < sel.size(); a++)
        for (size_t b = relvars_first_partner(sel, a); b < sel.size(); b++) {
            pa.push_back((int32_t)a); pb.push_back((int32_t)b); sv.push_back(S[a * sel.size() + b]);
        }
    std::vector<uint8_t> mark(sel.size(), 0);
    relvars_mark(sel, (int64_t)pa.size(), pa.data(), pb.data(), sv.data(), 1, gsize_u, cov_u, cutoff, mark);
    int n = 0;
    for (size_t a = 0; a < sel.size(); a++)
        if (mark[a]) vars[n++] = sel[a];
    vars[n] = -1;                                                        // 2483
    *n_vars = n;
    return RR_OK;
}

static int relvars_device_pairs(rr_packed *pk, const uint32_t *d_umask, const int32_t *gsize_u, const std::vector<int32_t> &sel, int cov_u,
                                double cutoff, std::vector<uint8_t> &mark, int64_t *pairs_tested)
{
    const int nsel = (int)sel.size();
    std::vector<int32_t> first((size_t)nsel);
    int64_t pairs = 0;
    for (int a = 0; a < nsel; a++) {
        first[a] = (int32_t)relvars_first_partner(sel, (size_t)a);
        pairs += nsel - first[a];
    }
    RR_CUDA(cudaSetDevice(pk->device));
    rr_alloc_stream(pk->st);
    dev_scope scope;
    constexpr unsigned UNSURE_CAP = 1u << 20;
    int32_t *d_sel = nullptr, *d_first = nullptr, *d_gu = nullptr;
    unsigned char *d_mark = nullptr;
    int4 *d_unsure = nullptr;
    unsigned int *d_count = nullptr, count = 0;
    int rc;
    if ((rc = scope.alloc(&d_sel, (size_t)nsel)) || (rc = scope.alloc(&d_first, (size_t)nsel)) || (rc = scope.alloc(&d_mark, (size_t)nsel)) ||
        (rc = scope.alloc(&d_unsure, (size_t)UNSURE_CAP)) || (rc = scope.alloc(&d_count, 1)) || (rc = scope.alloc(&d_gu, (size_t)5 * pk->N)))
        return rc;
    RR_CUDA(cudaMemcpyAsync(d_gu, gsize_u, sizeof(int32_t) * (size_t)5 * pk->N, cudaMemcpyHostToDevice, pk->st));
    RR_CUDA(cudaMemcpyAsync(d_sel, sel.data(), sizeof(int32_t) * (size_t)nsel, cudaMemcpyHostToDevice, pk->st));
    RR_CUDA(cudaMemcpyAsync(d_first, first.data(), sizeof(int32_t) * (size_t)nsel, cudaMemcpyHostToDevice, pk->st));
    RR_CUDA(cudaMemsetAsync(d_mark, 0, (size_t)nsel, pk->st));
    RR_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned int), pk->st));
    RR_CUDA(rr_launch_relvars_pairs(pk->d_bits, d_umask, pk->W32, d_sel, nsel, d_first, d_gu, cov_u, pk->d_lnfact, cutoff, d_mark, d_unsure,
                                    UNSURE_CAP, d_count, pk->st));
    RR_CUDA(cudaMemcpyAsync(mark.data(), d_mark, (size_t)nsel, cudaMemcpyDeviceToHost, pk->st));
    RR_CUDA(cudaMemcpyAsync(&count, d_count, sizeof(count), cudaMemcpyDeviceToHost, pk->st));
    RR_CUDA(cudaStreamSynchronize(pk->st));
    if (count > UNSURE_CAP) { rr_set_error("rr_relative_vars: %u undecided pairs exceed the list", count); return RR_E_NOMEM; }
    std::vector<int4> unsure(count);
    if (count) {
        RR_CUDA(cudaMemcpyAsync(unsure.data(), d_unsure, sizeof(int4) * count, cudaMemcpyDeviceToHost, pk->st));
        RR_CUDA(cudaStreamSynchronize(pk->st));
    }
    for (const int4 &u : unsure) {
        const double Z = rr_relative_score_host((uint32_t)u.z, (uint32_t)gsize_u[sel[u.y]], (uint32_t)gsize_u[sel[u.x]], (uint32_t)cov_u);
        if (Z > cutoff) mark[u.x] = mark[u.y] = 1;
    }
    if (pairs_tested) *pairs_tested = pairs;
    return RR_OK;
}

// Relative_Vars on the packed copy of the WHOLE MSA as it sits on the device after the scan - nothing is packed again; the part is a bitset over the packed row order, |G & U| comes from rr_k_masked_sizes and
// the all-pairs step from the tiled kernel with the mask ANDed into one operand.
extern "C" int rr_relative_vars_packed(rr_packed *pk, const int32_t *unterteilung, int u_no, const double *maxcorrs, double cutoff,
                                       int mingroup, int32_t *vars, int *n_vars, int64_t *pairs_tested)
{
    if (pairs_tested) *pairs_tested = 0;
    if (!pk || pk->phase != 3 || !unterteilung || !maxcorrs || !vars || !n_vars || mingroup < 1 || !(cutoff >= 0.0)) {
        rr_set_error("rr_relative_vars_packed: bad arguments");
        return RR_E_ARG;
    }
    vars[0] = -1;
    *n_vars = 0;
    const int R = pk->R, N = pk->N;
    if ((int)pk->h_perm.size() != R) { rr_set_error("rr_relative_vars_packed: the packed MSA carries no row order"); return RR_E_ARG; }
    std::vector<uint32_t> umask((size_t)pk->W32, 0u);
    int cov_u = 0;
    for (int rank = 0; rank < R; rank++)
        if (unterteilung[pk->h_perm[rank]] == u_no) { umask[rank >> 5] |= 1u << (rank & 31); cov_u++; }   // 2438, in packed row order
    if (cov_u < mingroup || N == 0) return RR_OK;
    RR_CUDA(cudaSetDevice(pk->device));
    rr_alloc_stream(pk->st);
    std::vector<int32_t> gsize_u((size_t)5 * N);
    std::vector<int32_t> sel;
    std::vector<uint8_t> mark;
    int rc;
    {
        dev_scope scope;
        uint32_t *d_umask = nullptr;
        int32_t *d_gu = nullptr;
        if ((rc = scope.alloc(&d_umask, umask.size())) || (rc = scope.alloc(&d_gu, gsize_u.size()))) return rc;
        RR_CUDA(cudaMemcpyAsync(d_umask, umask.data(), sizeof(uint32_t) * umask.size(), cudaMemcpyHostToDevice, pk->st));
        RR_CUDA(rr_launch_masked_sizes(pk->d_bits, d_umask, (int64_t)5 * N, pk->W32, d_gu, pk->st));
        RR_CUDA(cudaMemcpyAsync(gsize_u.data(), d_gu, sizeof(int32_t) * gsize_u.size(), cudaMemcpyDeviceToHost, pk->st));
        RR_CUDA(cudaStreamSynchronize(pk->st));
        relvars_select((int64_t)5 * N, maxcorrs, gsize_u.data(), cutoff, mingroup, sel);
        mark.assign(sel.size(), 0);
        if (!sel.empty() && (rc = relvars_device_pairs(pk, d_umask, gsize_u.data(), sel, cov_u, cutoff, mark, pairs_tested))) return rc;
    }
    int n = 0;
    for (size_t a = 0; a < sel.size(); a++)
        if (mark[a]) vars[n++] = sel[a];
    vars[n] = -1;
    *n_vars = n;
    return RR_OK;
}

extern "C" int rr_relative_vars(const rr_msa *msa, int device, const int32_t *unterteilung, int u_no, const double *maxcorrs,
                                double cutoff, int mingroup, int32_t *vars, int *n_vars, int64_t *pairs_tested)
{
    if (pairs_tested) *pairs_tested = 0;
    if (!msa || !unterteilung || !maxcorrs || !vars || !n_vars || mingroup < 1 || !(cutoff >= 0.0)) {
        rr_set_error("rr_relative_vars: bad arguments");
        return RR_E_ARG;
    }
    vars[0] = -1;
    *n_vars = 0;
    const int R = msa->rows, N = msa->cols;
    std::vector<int> rows;
    for (int r = 0; r < R; r++)
        if (unterteilung[r] == u_no) rows.push_back(r);                  // 2438
    const int cov_u = (int)rows.size();
    if (cov_u < mingroup || N == 0) return RR_OK;                        // no group can hold mingroup reads of the part (2449)
    // the part's rows as an MSA of their own
    rr_msa *sub = nullptr;
    int rc = rr_msa_alloc(cov_u, N, msa->codes, &sub);
    if (rc) return rc;
    for (int k = 0; k < cov_u; k++) memcpy(sub->cells + (size_t)k * N, rr_msa_row(msa, rows[k]), (size_t)N);
    rr_packed *pk = nullptr;
    rc = rr_pack(sub, device, &pk);
    rr_msa_free(sub);
    if (rc) return rc;
    const int32_t *gsize_u = pk->h_gsize.data();                         // |G & U|
    std::vector<int32_t> sel;
    relvars_select((int64_t)5 * N, maxcorrs, gsize_u, cutoff, mingroup, sel);
    std::vector<uint8_t> mark(sel.size(), 0);
    // the all-pairs step in one tiled kernel (rr_relvars.cu); pairs it cannot decide (score within 1e-9 of the cutoff)
    // come back in a list and are scored with the host libm
    if (!sel.empty()) rc = relvars_device_pairs(pk, nullptr, gsize_u, sel, cov_u, cutoff, mark, pairs_tested);
    rr_packed_free(pk);
    if (rc) return rc;
    int n = 0;
    for (size_t a = 0; a < sel.size(); a++)
        if (mark[a]) vars[n++] = sel[a];
    vars[n] = -1;
    *n_vars = n;
    return RR_OK;
}

// ---------------------------------------------------------------------------------------
// Kmeans (RepeatResolver.c:2604-2821): signatures and the dissolution of small clusters on the host, the two read x read
// sweeps and the centroids on the device (rr_kmeans.cu).  The host pieces and the integer rules shared with the kernels
// (rr_kmeans.h) are pinned against the unmodified reference on the CPU, the device path on a B200 (tests/test_zz_gpu_kmeans.py).
// ---------------------------------------------------------------------------------------
static inline int host_class(uint8_t c, int codes)                       // 304-329, as rr_classify in rr_pack.cu
{
    if (codes) return c < 5 ? (int)c : 5;
    switch (c) {
    case 'a': case 'A': return 0;
    case 'c': case 'C': return 1;
    case 'g': case 'G': return 2;
    case 't': case 'T': return 3;
    case '-': case '_': return 4;
    default: return 5;
    }
}

// 2616-2642: the reads of the part and their signatures over the selected groups
extern "C" int rr_kmeans_signatures(const rr_msa *msa, const int32_t *unterteilung, int u_no, const int32_t *vars, int n_vars,
                                    int32_t *reads_out, int *anzahl_out, uint64_t *sig_out)
{
    if (!msa || !unterteilung || n_vars < 0 || (n_vars && !vars) || !reads_out || !anzahl_out) {
        rr_set_error("rr_kmeans_signatures: bad arguments");
        return RR_E_ARG;
    }
    for (int j = 0; j < n_vars; j++)
        if (vars[j] < 0 || vars[j] >= 5 * msa->cols) { rr_set_error("rr_kmeans_signatures: group %d out of range", vars[j]); return RR_E_ARG; }
    const int scv = n_vars / 64 + 1;                                      // 2626
    int anzahl = 0;
    for (int r = 0; r < msa->rows; r++)
        if (unterteilung[r] == u_no) reads_out[anzahl++] = r;
    *anzahl_out = anzahl;
    if (!sig_out) return RR_OK;
    std::fill(sig_out, sig_out + (size_t)anzahl * scv, (uint64_t)0);
    // site and symbol of every selected group once, then the reads in parallel (one read = one row of the matrix)
    std::vector<int32_t> vsite((size_t)n_vars);
    std::vector<uint8_t> vsym((size_t)n_vars);
    for (int j = 0; j < n_vars; j++) { vsite[j] = vars[j] / 5; vsym[j] = (uint8_t)(vars[j] % 5); }
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)16, (int64_t)std::thread::hardware_concurrency(),
                                                                  (int64_t)anzahl * n_vars / 2000000 + 1}));
    uint8_t cls[256];                                                    // the cell's class once per byte value, no branch per group
    for (int c = 0; c < 256; c++) cls[c] = (uint8_t)host_class((uint8_t)c, msa->codes);
    auto work = [&](int t) {
        for (int i = (int)((int64_t)anzahl * t / nt); i < (int)((int64_t)anzahl * (t + 1) / nt); i++) {
            const uint8_t *row = rr_msa_row(msa, reads_out[i]);
            uint64_t *sg = sig_out + (size_t)i * scv;
            for (int j0 = 0; j0 < n_vars; j0 += 64) {                    // one signature word in a register
                const int n = std::min(64, n_vars - j0);
                uint64_t w = 0;
                for (int b = 0; b < n; b++) w |= (uint64_t)(cls[row[vsite[j0 + b]]] == vsym[j0 + b]) << b;
                sg[j0 / 64] = w;
            }
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    return RR_OK;
}

// 2728-2757 and 2790-2797: clusters of at most `min` reads are dissolved into clusters of at least `min`, in read order,
// sizes updated as it goes; returns the number of non-empty clusters
// the clusters that can ever hold min >= 2 reads during the dissolution (2727-2752): a cluster only grows when a displaced
// read picks it, and a read picks an admissible cluster (already >= min) or, when nothing scores above 0, cluster 0
static void km_candidate_clusters(int anzahl, const int32_t *cluster_in, std::vector<int32_t> &J)
{
    std::vector<int> size((size_t)anzahl + 1, 0);
    for (int i = 0; i < anzahl; i++)
        if (cluster_in[i] >= 0 && cluster_in[i] < anzahl) size[cluster_in[i]]++;
    J.clear();
    for (int j = 0; j < anzahl; j++)
        if (j == 0 || size[j] >= 2) J.push_back(j);
}

// J / S: optional table of the scores S[i * nJ + k] = GrMatch(Centroids[J[k]], VarSigs[i]) for the clusters of
// km_candidate_clusters (ascending), made on the device by rr_kmeans; without it the scores are computed here
static int km_finish_impl(int anzahl, int scv, const uint64_t *sig, const uint64_t *cen, const int32_t *cluster_in, int mingroup,
                          int32_t *cluster_out, int *n_clusters, const int32_t *J, int nJ, const int32_t *S)
{
    if (anzahl < 0 || scv < 1 || !n_clusters || (anzahl && (!sig || !cen || !cluster_in || !cluster_out))) {
        rr_set_error("rr_kmeans_finish: bad arguments");
        return RR_E_ARG;
    }
    std::vector<int> size((size_t)anzahl + 1, 0);
    for (int i = 0; i < anzahl; i++) {
        if (cluster_in[i] < 0 || cluster_in[i] >= anzahl) { rr_set_error("rr_kmeans_finish: cluster %d out of range", cluster_in[i]); return RR_E_ARG; }
        cluster_out[i] = cluster_in[i];
        size[cluster_in[i]]++;
    }
    // the reads are visited in order and the sizes change as they go (that order is part of the reference's result); the
    // search for one read's best cluster is independent work over j: (clusters of at least min reads) x scv word
    // operations.  Usually that is a few dozen clusters and cheaper than starting threads (measured: 766 reads with 70 k
    // groups, 120 clusters: 176 ms with five threads started per displaced read, 9 ms serial), so threads are only used where
    // one search is worth more than a millisecond.
    const int nt_max = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)16, (int64_t)std::thread::hardware_concurrency()));
    std::vector<int> tb((size_t)nt_max), tj((size_t)nt_max);
    for (int min = 2; min < mingroup; min++) {
        int64_t admissible = 0;                                          // at the start of the level; later moves change it little
        for (int j = 0; j < anzahl; j++) admissible += size[j] >= min;
        const int nt = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)nt_max, admissible * scv / 4000000));
        for (int i = 0; i < anzahl; i++)
            if (size[cluster_out[i]] <= min && S) {                      // scores from the table: the same loop over ascending j
                int best = 0, best_j = 0;
                const int32_t *row = S + (size_t)i * nJ;
                for (int k = 0; k < nJ; k++) {
                    const int j = J[k];
                    if (size[j] >= min && cluster_out[i] != j && row[k] > best && i != j) { best = row[k]; best_j = j; }
                }
                size[cluster_out[i]]--;
                cluster_out[i] = best_j;
                size[best_j]++;
            } else if (size[cluster_out[i]] <= min) {
                auto search = [&](int t) {
                    int best = 0, best_j = 0;                            // first best in ascending j, as the serial loop
                    for (int j = (int)((int64_t)anzahl * t / nt); j < (int)((int64_t)anzahl * (t + 1) / nt); j++)
                        if (size[j] >= min && cluster_out[i] != j) {
                            const int score = rr_km_match(cen + (size_t)j * scv, sig + (size_t)i * scv, scv);
                            if (score > best && i != j) { best = score; best_j = j; }
                        }
                    tb[t] = best; tj[t] = best_j;
                };
                if (nt == 1) search(0);
                else {
                    std::vector<std::thread> th;
                    for (int t = 0; t < nt; t++) th.emplace_back(search, t);
                    for (auto &t : th) t.join();
                }
                int best = 0, best_j = 0;
                for (int t = 0; t < nt; t++)
                    if (tb[t] > best) { best = tb[t]; best_j = tj[t]; }  // strict >: the earliest slice wins ties
                size[cluster_out[i]]--;
                cluster_out[i] = best_j;
                size[best_j]++;
            }
    }
    int n = 0;
    for (int i = 0; i < anzahl; i++) n += size[i] > 0;
    *n_clusters = n;
    return RR_OK;
}

extern "C" int rr_kmeans_finish(int anzahl, int scv, const uint64_t *sig, const uint64_t *cen, const int32_t *cluster_in, int mingroup,
                                int32_t *cluster_out, int *n_clusters)
{
    return km_finish_impl(anzahl, scv, sig, cen, cluster_in, mingroup, cluster_out, n_clusters, nullptr, 0, nullptr);
}

// test hook: the dissolution with the score table rr_kmeans makes on the device, here filled on the host
extern "C" int rr_debug_kmeans_finish_table(int anzahl, int scv, const uint64_t *sig, const uint64_t *cen, const int32_t *cluster_in, int mingroup,
                                      int32_t *cluster_out, int *n_clusters)
{
    if (anzahl < 0 || scv < 1 || (anzahl && (!sig || !cen || !cluster_in))) { rr_set_error("rr_debug_kmeans_finish_table: bad arguments"); return RR_E_ARG; }
    std::vector<int32_t> J;
    km_candidate_clusters(anzahl, cluster_in, J);
    std::vector<int32_t> S((size_t)anzahl * J.size());
    for (int i = 0; i < anzahl; i++)
        for (size_t k = 0; k < J.size(); k++) S[(size_t)i * J.size() + k] = rr_km_match(cen + (size_t)J[k] * scv, sig + (size_t)i * scv, scv);
    return km_finish_impl(anzahl, scv, sig, cen, cluster_in, mingroup, cluster_out, n_clusters, J.data(), (int)J.size(), S.data());
}

// function-level hooks on the rules the kernels share with the host (rr_kmeans.h)
extern "C" int rr_kmeans_top5_host(int anzahl, int scv, const uint64_t *sig, int i, int32_t *best_j)
{
    if (anzahl < 1 || scv < 1 || !sig || i < 0 || i >= anzahl || !best_j) return RR_E_ARG;
    int bs[5] = {0, 0, 0, 0, 0}, bj[5] = {0, 0, 0, 0, 0};
    for (int j = 0; j < anzahl; j++) rr_km_top5_step(bs, bj, rr_km_match(sig + (size_t)j * scv, sig + (size_t)i * scv, scv), j);
    for (int k = 0; k < 5; k++) best_j[k] = bj[k];
    return RR_OK;
}

extern "C" uint64_t rr_kmeans_majority5_host(uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t e)
{
    return rr_km_majority5(a, b, c, d, e);
}

extern "C" int rr_kmeans(const rr_msa *msa, int device, int32_t *unterteilung, int u_no, const int32_t *vars, int n_vars,
                         int mingroup, int *n_clusters)
{
    if (!msa || !unterteilung || !n_clusters || n_vars < 0 || (n_vars && !vars)) { rr_set_error("rr_kmeans: bad arguments"); return RR_E_ARG; }
    *n_clusters = 0;
    const int ndev = rr_device_count();
    if (ndev <= 0) { rr_set_error("no CUDA device: Kmeans has no CPU fallback"); return RR_E_NODEV; }
    if (device < 0 || device >= ndev) { rr_set_error("device %d out of range (%d devices)", device, ndev); return RR_E_ARG; }
    const int scv = n_vars / 64 + 1;
    std::vector<int32_t> reads((size_t)std::max(msa->rows, 1));
    int anzahl = 0, rc;
    if ((rc = rr_kmeans_signatures(msa, unterteilung, u_no, vars, n_vars, reads.data(), &anzahl, nullptr))) return rc;   // validates; the reads of the part
    if (anzahl == 0) return RR_OK;
    std::vector<uint64_t> sig((size_t)anzahl * scv), cen((size_t)anzahl * scv);
    std::vector<int32_t> cluster((size_t)anzahl), final_cluster((size_t)anzahl), J, S;
    RR_CUDA(cudaSetDevice(device));
    cudaStream_t st = nullptr;
    RR_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    rr_alloc_stream(st);
    {
        dev_scope scope;
        uint64_t *d_sig = nullptr, *d_cen = nullptr;
        int32_t *d_best = nullptr, *d_cluster = nullptr, *d_vars = nullptr, *d_panel = nullptr, *d_state = nullptr;
        uint8_t *d_rows = nullptr;
        // scores of the two sweeps: panels of rows j, all reads i each; 64 MB at most
        const int panel_rows = (int)std::max<size_t>(1, std::min<size_t>((size_t)anzahl, ((size_t)16 << 20) / (size_t)anzahl));
        cudaError_t e = cudaSuccess;
        // the part's rows go to the device in slices of at most 256 MB; their signatures are made there (the host loop over
        // reads x groups was most of the call: 5e7 byte classifications for a part of 766 reads and 70 k groups)
        const size_t N = (size_t)msa->cols;
        const int slice = (int)std::max<size_t>(1, std::min<size_t>((size_t)anzahl, ((size_t)256 << 20) / std::max<size_t>(N, 1)));
        if ((rc = scope.alloc(&d_sig, sig.size())) || (rc = scope.alloc(&d_cen, cen.size())) || (rc = scope.alloc(&d_best, (size_t)anzahl * 5)) ||
            (rc = scope.alloc(&d_cluster, (size_t)anzahl)) || (rc = scope.alloc(&d_vars, (size_t)std::max(n_vars, 1))) ||
            (rc = scope.alloc(&d_rows, (size_t)slice * std::max<size_t>(N, 1))) || (rc = scope.alloc(&d_panel, (size_t)panel_rows * anzahl)) ||
            (rc = scope.alloc(&d_state, (size_t)6 * anzahl))) {
            // fall through to the clean-up below
        } else {
            if (n_vars) e = cudaMemcpyAsync(d_vars, vars, sizeof(int32_t) * (size_t)n_vars, cudaMemcpyHostToDevice, st);
            for (int i0 = 0; i0 < anzahl && e == cudaSuccess; i0 += slice) {
                const int n = std::min(slice, anzahl - i0);
                for (int i = i0; i < i0 + n && e == cudaSuccess;) {      // runs of consecutive rows of a matrix in one copy
                    int k = i + 1;
                    while (msa->cells && k < i0 + n && reads[k] == reads[k - 1] + 1) k++;
                    e = cudaMemcpyAsync(d_rows + (size_t)(i - i0) * N, rr_msa_row(msa, reads[i]), (size_t)(k - i) * N, cudaMemcpyHostToDevice, st);
                    i = k;
                }
                if (e == cudaSuccess)
                    e = rr_launch_kmeans_signatures(d_rows, msa->cols, msa->codes, d_vars, n_vars, n, scv, d_sig + (size_t)i0 * scv, st);
            }
            if (e != cudaSuccess ||
                (e = cudaMemcpyAsync(sig.data(), d_sig, sizeof(uint64_t) * sig.size(), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
                (e = rr_launch_kmeans_sweeps(d_sig, anzahl, scv, d_best, d_cen, d_cluster, d_panel, panel_rows, d_state, st)) != cudaSuccess ||
                (e = cudaMemcpyAsync(cen.data(), d_cen, sizeof(uint64_t) * cen.size(), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
                (e = cudaMemcpyAsync(cluster.data(), d_cluster, sizeof(int32_t) * (size_t)anzahl, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
                (e = cudaStreamSynchronize(st)) != cudaSuccess) {
                rr_set_error("CUDA error %s in rr_kmeans", cudaGetErrorString(e));
                rc = RR_E_CUDA;
            }
            // the scores of the dissolution (reads x clusters that can ever be admissible) while signatures and centroids are
            // still on the device: on the host that search was most of what remained of the call
            if (!rc) {
                km_candidate_clusters(anzahl, cluster.data(), J);
                int32_t *d_J = nullptr, *d_S = nullptr;
                if ((size_t)anzahl * J.size() <= ((size_t)1 << 28) && !scope.alloc(&d_J, J.size()) && !scope.alloc(&d_S, (size_t)anzahl * J.size())) {
                    S.resize((size_t)anzahl * J.size());
                    if ((e = cudaMemcpyAsync(d_J, J.data(), sizeof(int32_t) * J.size(), cudaMemcpyHostToDevice, st)) != cudaSuccess ||
                        (e = rr_launch_kmeans_scores(d_sig, d_cen, d_J, (int)J.size(), anzahl, scv, d_S, st)) != cudaSuccess ||
                        (e = cudaMemcpyAsync(S.data(), d_S, sizeof(int32_t) * S.size(), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
                        (e = cudaStreamSynchronize(st)) != cudaSuccess) {
                        rr_set_error("CUDA error %s in rr_kmeans", cudaGetErrorString(e));
                        rc = RR_E_CUDA;
                    }
                } else {
                    cudaGetLastError();
                    S.clear();                                           // too large a table: the host computes the scores as it goes
                }
            }
        }
    }   // device buffers go back to the pool while the stream still exists
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    rr_alloc_stream(nullptr);
    if (rc) return rc;
    if ((rc = km_finish_impl(anzahl, scv, sig.data(), cen.data(), cluster.data(), mingroup, final_cluster.data(), n_clusters,
                             S.empty() ? nullptr : J.data(), (int)J.size(), S.empty() ? nullptr : S.data())))
        return rc;
    int max_u = 0;                                                       // 2814-2815
    for (int r = 0; r < msa->rows; r++) max_u = std::max(max_u, unterteilung[r]);
    for (int i = 0; i < anzahl; i++) unterteilung[reads[i]] = final_cluster[i] + max_u + 1;
    return RR_OK;
}

// ---------------------------------------------------------------------------------------
// scan
// ---------------------------------------------------------------------------------------
template <typename T>
static int upload(T **d, const std::vector<T> &h, cudaStream_t st)
{
    int rc = dev_alloc(d, h.size());
    if (rc) return rc;
    if (!h.empty()) RR_CUDA(cudaMemcpyAsync(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, st));
    return RR_OK;
}


extern "C" int rr_scan(rr_packed *pk, const rr_scan_opts *opts, rr_scan_stats *stats)
{
    if (!pk || !opts) { rr_set_error("rr_scan: bad arguments"); return RR_E_ARG; }
    if (opts->part_count < 1 || opts->part_index < 0 || opts->part_index >= opts->part_count) {
        rr_set_error("rr_scan: part %d of %d", opts->part_index, opts->part_count);
        return RR_E_ARG;
    }
    if (pk->phase != 3) { rr_set_error("rr_scan: the packed MSA is incomplete (rr_pack_rows without rr_pack_set_spans / rr_pack_finish)"); return RR_E_ARG; }
    if (opts->flags & RR_FLAG_HOST_FINALIZE) {
        rr_set_error("rr_scan: RR_FLAG_HOST_FINALIZE is a flag of rr_maxcorr_run; after rr_scan + rr_scan_fetch call rr_scan_finalize");
        return RR_E_ARG;
    }
    RR_CUDA(cudaSetDevice(pk->device));
    rr_alloc_stream(pk->st);
    const int R = pk->R, N = pk->N, mincov = opts->mincov;
    event_scope ev;                                                      // destroyed on every exit path
    cudaEvent_t e0, e1, e2;
    if (ev.create(&e0) || ev.create(&e1) || ev.create(&e2)) { rr_set_error("cudaEventCreate failed"); return RR_E_CUDA; }
    RR_CUDA(cudaEventRecord(e0, pk->st));

    int variant = opts->variant;
    if (variant == RR_VARIANT_AUTO) variant = rr_umma_available() ? RR_VARIANT_UMMA_MXF4 : RR_VARIANT_BITSET;
    if (variant != RR_VARIANT_BITSET && variant != RR_VARIANT_UMMA && variant != RR_VARIANT_UMMA_F4 && variant != RR_VARIANT_UMMA_MXF4) { rr_set_error("unknown variant %d", variant); return RR_E_ARG; }

    // the accumulators hold 4 x count (operand elements 0 / 2); fp32 accumulation is exact only below 2^24, i.e. 2^22
    // reads: deeper MSAs use the int8/int32 coding
    if ((variant == RR_VARIANT_UMMA_MXF4 || variant == RR_VARIANT_UMMA_F4) && R >= (1 << 22)) variant = RR_VARIANT_UMMA;
    const int umma_mode = variant == RR_VARIANT_UMMA_MXF4 ? 2 : variant == RR_VARIANT_UMMA_F4 ? 1 : 0;
    // ---- host plan: filters, first-break columns, tiles, partition (O(N)); cached between scans ----
    const bool general = !pk->contiguous || (opts->flags & RR_FLAG_GENERAL_BREAK);
    scan_cache &C = pk->cache;
    int rc;
    if (!(C.valid && C.mincov == mincov && C.variant == variant && C.part_index == opts->part_index &&
          C.part_count == opts->part_count && C.general == general)) {
        C.valid = false;
        C.sb.release();
        C.plan = rr_plan();
        std::vector<int32_t> breakcol(N);
        if (general) {
            if ((rc = dev_alloc(&C.sb.breakcol, (size_t)N))) return rc;
            RR_CUDA(rr_launch_general_break(pk->d_covbits, pk->W32, N, mincov, C.sb.breakcol, pk->st));
            if (N) RR_CUDA(cudaMemcpyAsync(breakcol.data(), C.sb.breakcol, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, pk->st));
            RR_CUDA(cudaStreamSynchronize(pk->st));
        } else {
            rc = rr_breakcols_from_spans(pk->h_start.data(), pk->h_end.data(), R, N, mincov, breakcol.data());
            if (rc) return rc;
        }
        RR_TRACE("breakcols");
        const int ti = variant == RR_VARIANT_BITSET ? rr_bitset_ti() : rr_umma_row_sites();
        const int tj = variant == RR_VARIANT_BITSET ? rr_bitset_tj() : rr_umma_col_sites();
        rr_plan_build(C.plan, R, N, mincov, pk->h_gsize.data(), pk->h_coverage.data(), breakcol.data(),
                      pk->contiguous && !general ? pk->h_start.data() : nullptr,
                      pk->contiguous && !general ? pk->h_end.data() : nullptr, pk->class_start, pk->n_classes, ti, tj, variant == RR_VARIANT_BITSET ? 32 : rr_umma_kblock(umma_mode),
                      // epilogue of a tile in K-block equivalents / overlap: 8.3e-5 ms per tile against 3.4e-6 (mxf4),
                      // 2.4e-6 (e2m1) and 4.2e-6 ms (int8) per K block, measured part by part at config 2
                      variant == RR_VARIANT_BITSET ? 8 : umma_mode == 2 ? 25 : umma_mode == 1 ? 34 : 20,
                      variant == RR_VARIANT_BITSET ? 0 : 80, opts->part_index, opts->part_count);
        RR_TRACE("plan");
        if ((rc = upload(&C.sb.rowok, C.plan.rowok, pk->st))) return rc;
        if ((rc = upload(&C.sb.colok, C.plan.colok, pk->st))) return rc;
        if (!general && (rc = upload(&C.sb.breakcol, breakcol, pk->st))) return rc;
        if ((rc = upload(&C.sb.rowsites, C.plan.rowsites, pk->st))) return rc;
        if ((rc = upload(&C.sb.unit_prefix, C.plan.unit_prefix, pk->st))) return rc;
        if ((rc = upload(&C.sb.unit_cb0, C.plan.unit_cb0, pk->st))) return rc;
        if ((rc = upload(&C.sb.word_hi, C.plan.k_hi, pk->st))) return rc;
        if ((rc = upload(&C.sb.word_lo, C.plan.k_lo, pk->st))) return rc;
        RR_CUDA(cudaStreamSynchronize(pk->st));  // the host vectors above are locals
        C.mincov = mincov; C.variant = variant; C.part_index = opts->part_index; C.part_count = opts->part_count;
        C.general = general; C.plan_id = pk->next_plan_id++;
        C.valid = true;
    }
    rr_plan &plan = C.plan;
    scan_buffers &sb = C.sb;

    // running maxima: 0.0 / no partner
    if (!(opts->flags & RR_FLAG_SKIP_SEED)) RR_CUDA(rr_launch_init_best(pk->d_best, (int64_t)5 * N, pk->st));
    RR_CUDA(cudaMemsetAsync(pk->d_counters, 0, sizeof(unsigned long long) * 8, pk->st));
    RR_TRACE("uploads+init");

    rr_scan_params P;
    memset(&P, 0, sizeof P);
    P.R = R; P.N = N; P.W32 = pk->W32; P.mincov = mincov; P.flags = opts->flags;
    P.bits = pk->d_bits; P.gsize = pk->d_gsize; P.rowok = sb.rowok; P.colok = sb.colok;
    P.breakcol = sb.breakcol; P.rowsites = sb.rowsites; P.n_rowsites = plan.n_rowsites;
    P.lnfact = pk->d_lnfact; P.best = pk->d_best; P.counters = pk->d_counters;
    P.unit_prefix = sb.unit_prefix; P.unit_cb0 = sb.unit_cb0; P.n_rowblocks = plan.n_rowblocks;
    P.rb_lo = plan.rb_lo; P.rb_hi = plan.rb_hi; P.word_hi = sb.word_hi; P.word_lo = sb.word_lo;
    P.n_colblocks = plan.n_colblocks;
    P.n_classes = plan.n_classes;

    if (variant != RR_VARIANT_BITSET && !(opts->flags & (RR_FLAG_NO_PRUNE | RR_DEBUG_MMA_ONLY))) {
        // deferred exact evaluation: room for one candidate per ~2 500 pair tests of this part (config 2: one per 6 000
        // survives tier 2), 20 bytes each; what does not fit is evaluated in place by the scan kernel
        size_t want = (size_t)std::min<int64_t>(std::max<int64_t>(plan.part_pairs / 2500, (int64_t)1 << 20), (int64_t)96 << 20);
        if (g_deferred_cap) want = (size_t)g_deferred_cap;               // test hook: a list that overflows
        if (want > pk->deferred_cap) {
            rr_dev_free(pk->d_deferred);
            pk->d_deferred = nullptr; pk->deferred_cap = 0;
            if (rr_dev_malloc((void **)&pk->d_deferred, want * sizeof(rr_cand_p)) == cudaSuccess) pk->deferred_cap = want;
            else cudaGetLastError();       // no list: the kernel evaluates in place
        }
        P.deferred = pk->d_deferred;
        P.deferred_cap = pk->d_deferred ? std::min(pk->deferred_cap, want) : 0;
        P.defer_mode = 1;
    }

    RR_CUDA(cudaEventRecord(e1, pk->st));
    int64_t executed = 0;
    if (plan.rb_hi > plan.rb_lo && plan.unit_prefix[plan.rb_hi] > plan.unit_prefix[plan.rb_lo]) {
        if (variant == RR_VARIANT_BITSET) {
            if (!(opts->flags & RR_FLAG_SEED_ONLY)) RR_CUDA(rr_launch_scan_bitset(P, pk->n_sm, pk->st));  // no seeding pass in this variant
        } else {
            rc = rr_umma_scan(pk->umma, umma_mode, C.plan_id, P, plan, pk->n_sm, pk->st);
            if (rc) return rc;
        }
    }
    RR_TRACE("launched");
    RR_CUDA(cudaEventRecord(e2, pk->st));
    unsigned long long counters[8];
    RR_CUDA(cudaMemcpyAsync(counters, pk->d_counters, sizeof counters, cudaMemcpyDeviceToHost, pk->st));
    RR_CUDA(cudaStreamSynchronize(pk->st));
    RR_CUDA(cudaGetLastError());
    RR_TRACE("kernels done");
    pk->have_result = true;
    executed = plan.executed_ops;

    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->pair_tests = (int64_t)counters[0];
        stats->exact_evals = (int64_t)counters[1];
        stats->bound_evals = (int64_t)(counters[2] + counters[4]);  // tier-1 table bounds (bitset kernel) + tier-2 evaluations (UMMA)
        stats->work_units = (int64_t)counters[3];
        stats->executed_ops = executed;
        stats->variant = variant;
        stats->rows = R; stats->cols = N; stats->row_sites = plan.n_rowsites;
        stats->general_break = general ? 1 : 0;
        stats->h2d_ms = pk->h2d_ms; stats->pack_ms = pk->pack_ms;
        cudaEventElapsedTime(&stats->prepare_ms, e0, e1);
        cudaEventElapsedTime(&stats->kernel_ms, e1, e2);
    }
    if ((int64_t)counters[0] != plan.part_pairs && !(opts->flags & (RR_DEBUG_MMA_ONLY | RR_FLAG_SEED_ONLY))) {
        rr_set_error("pair-test count mismatch: device %llu, host plan %lld", counters[0], (long long)plan.part_pairs);
        return RR_E_CUDA;
    }
    return RR_OK;
}

extern "C" int rr_scan_set_thresholds(rr_packed *pk, const double *thr)
{
    if (!pk || (!thr && pk->N > 0)) { rr_set_error("rr_scan_set_thresholds: bad arguments"); return RR_E_ARG; }
    if (!pk->have_result) { rr_set_error("rr_scan_set_thresholds before a seeding scan"); return RR_E_ARG; }
    RR_CUDA(cudaSetDevice(pk->device));
    rr_alloc_stream(pk->st);
    const size_t G = (size_t)5 * pk->N;
    if (G == 0) return RR_OK;
    double *d_thr = nullptr;
    int rc = dev_alloc(&d_thr, G);
    if (rc) return rc;
    RR_CUDA(cudaMemcpyAsync(d_thr, thr, sizeof(double) * G, cudaMemcpyHostToDevice, pk->st));
    RR_CUDA(rr_launch_raise_best(pk->d_best, d_thr, (int64_t)G, pk->st));
    RR_CUDA(cudaStreamSynchronize(pk->st));
    rr_dev_free(d_thr);
    return RR_OK;
}

extern "C" int rr_scan_values_device(rr_packed *pk, double *d_values)
{
    if (!pk || (!d_values && pk->N > 0)) { rr_set_error("rr_scan_values_device: bad arguments"); return RR_E_ARG; }
    if (!pk->have_result) { rr_set_error("rr_scan_values_device before rr_scan"); return RR_E_ARG; }
    RR_CUDA(cudaSetDevice(pk->device));
    RR_CUDA(rr_launch_best_values(pk->d_best, d_values, (int64_t)5 * pk->N, pk->st));
    RR_CUDA(cudaStreamSynchronize(pk->st));
    return RR_OK;
}

extern "C" int rr_scan_set_thresholds_device(rr_packed *pk, const double *d_thr)
{
    if (!pk || (!d_thr && pk->N > 0)) { rr_set_error("rr_scan_set_thresholds_device: bad arguments"); return RR_E_ARG; }
    if (!pk->have_result) { rr_set_error("rr_scan_set_thresholds_device before a seeding scan"); return RR_E_ARG; }
    RR_CUDA(cudaSetDevice(pk->device));
    RR_CUDA(rr_launch_raise_best(pk->d_best, d_thr, (int64_t)5 * pk->N, pk->st));
    RR_CUDA(cudaStreamSynchronize(pk->st));
    return RR_OK;
}

extern "C" int rr_scan_fetch(rr_packed *pk, double *maxcorr, int32_t *argmax)
{
    if (!pk || (!maxcorr && pk->N > 0)) { rr_set_error("rr_scan_fetch: bad arguments"); return RR_E_ARG; }
    if (!pk->have_result) { rr_set_error("rr_scan_fetch before rr_scan"); return RR_E_ARG; }
    RR_CUDA(cudaSetDevice(pk->device));
    const size_t G = (size_t)5 * pk->N;
    std::vector<rr_best_t> best(G);
    if (G) RR_CUDA(cudaMemcpyAsync(best.data(), pk->d_best, sizeof(rr_best_t) * G, cudaMemcpyDeviceToHost, pk->st));
    RR_CUDA(cudaStreamSynchronize(pk->st));
    for (size_t g = 0; g < G; g++) {
        double z;
        memcpy(&z, &best[g].z, sizeof z);
        maxcorr[g] = z;
        if (argmax) argmax[g] = best[g].p == ~0ull ? -1 : (int32_t)best[g].p;
    }
    return RR_OK;
}

// RR_FLAG_HOST_FINALIZE as a call of its own (rr_maxcorr_run uses it after its merge; one-process-per-GPU callers after
// theirs): maxcorr[g] of every group with a partner is re-evaluated with the host libm from the pair's counts, which
// come from the device bitsets (rr_pair_counts) - so the "%f" text is byte-identical to the reference's.
extern "C" int rr_scan_finalize(rr_packed *pk, double *maxcorr, const int32_t *argmax)
{
    if (!pk || ((!maxcorr || !argmax) && pk->N > 0)) { rr_set_error("rr_scan_finalize: bad arguments"); return RR_E_ARG; }
    const size_t G = (size_t)5 * pk->N;
    std::vector<int32_t> gi, gj;
    std::vector<size_t> idx;
    for (size_t g = 0; g < G; g++)
        if (argmax[g] >= 0) {
            if ((size_t)argmax[g] >= G) { rr_set_error("rr_scan_finalize: partner %d out of range", argmax[g]); return RR_E_ARG; }
            const int32_t a = (int32_t)g, b = argmax[g];
            gi.push_back(std::min(a, b)); gj.push_back(std::max(a, b)); idx.push_back(g);
        }
    std::vector<int32_t> cnt(4 * gi.size());
    int rc = rr_pair_counts(pk, (int64_t)gi.size(), gi.data(), gj.data(), cnt.data());
    if (rc) return rc;
    const std::vector<int32_t> &gs = pk->h_gsize;
    const int nt = (int)std::max<size_t>(1, std::min<size_t>({(size_t)8, (size_t)std::thread::hardware_concurrency(), idx.size() / 4096 + 1}));
    auto work = [&](int t) {
        for (size_t k = idx.size() * t / nt; k < idx.size() * (t + 1) / nt; k++)
            maxcorr[idx[k]] = rr_score_host((uint32_t)cnt[4 * k], (uint32_t)cnt[4 * k + 1], (uint32_t)cnt[4 * k + 2],
                                            (uint32_t)cnt[4 * k + 3], gs[gi[k]], gs[gj[k]]);
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    return RR_OK;
}

// ---------------------------------------------------------------------------------------
// test hook (include/rr_debug.h): one accumulator tile of the tcgen05 kernel
// ---------------------------------------------------------------------------------------
extern "C" int rr_debug_umma_counts(rr_packed *pk, const rr_scan_opts *opts, int row_tile, int col_tile, int32_t *counts,
                                    int32_t *row_groups, int32_t *col_groups, int *n_row_tiles, int *n_col_tiles)
{
    if (!pk || !opts) { rr_set_error("rr_debug_umma_counts: bad arguments"); return RR_E_ARG; }
    scan_cache &C = pk->cache;
    int variant = opts->variant == RR_VARIANT_AUTO ? RR_VARIANT_UMMA_MXF4 : opts->variant;
    if (!C.valid || !pk->umma || C.variant != variant || C.mincov != opts->mincov || C.part_index != opts->part_index ||
        C.part_count != opts->part_count || variant == RR_VARIANT_BITSET) {
        rr_set_error("rr_debug_umma_counts: call rr_scan with the same options (a tcgen05 variant) first");
        return RR_E_ARG;
    }
    if (n_row_tiles) *n_row_tiles = C.plan.n_rowblocks;
    if (n_col_tiles) *n_col_tiles = C.plan.n_colblocks;
    if (!counts) return RR_OK;
    RR_CUDA(cudaSetDevice(pk->device));
    rr_alloc_stream(pk->st);
    const int mode = variant == RR_VARIANT_UMMA_MXF4 ? 2 : variant == RR_VARIANT_UMMA_F4 ? 1 : 0;
    const int M = 128, NC = 5 * rr_umma_col_sites(), RS = rr_umma_row_sites();
    dev_scope scope;
    int32_t *d_out = nullptr;
    unsigned long long *d_cnt = nullptr;
    int rc;
    if ((rc = scope.alloc(&d_out, (size_t)M * NC)) || (rc = scope.alloc(&d_cnt, 8))) return rc;
    RR_CUDA(cudaMemsetAsync(d_cnt, 0, 8 * sizeof(unsigned long long), pk->st));
    rr_scan_params P;
    memset(&P, 0, sizeof P);
    scan_buffers &sb = C.sb;
    P.R = pk->R; P.N = pk->N; P.W32 = pk->W32; P.mincov = opts->mincov; P.flags = 0;
    P.bits = pk->d_bits; P.gsize = pk->d_gsize; P.rowok = sb.rowok; P.colok = sb.colok;
    P.breakcol = sb.breakcol; P.rowsites = sb.rowsites; P.n_rowsites = C.plan.n_rowsites;
    P.lnfact = pk->d_lnfact; P.best = pk->d_best; P.counters = d_cnt;
    P.n_rowblocks = C.plan.n_rowblocks; P.n_colblocks = C.plan.n_colblocks; P.n_classes = C.plan.n_classes;
    if ((rc = rr_umma_dump_tile(pk->umma, mode, C.plan_id, P, C.plan, row_tile, col_tile, d_out, pk->n_sm, pk->st))) return rc;
    RR_CUDA(cudaMemcpyAsync(counts, d_out, sizeof(int32_t) * (size_t)M * NC, cudaMemcpyDeviceToHost, pk->st));
    RR_CUDA(cudaStreamSynchronize(pk->st));
    if (row_groups)
        for (int r = 0; r < M; r++) {
            const int l = r & 31, slab = r >> 5;
            const int site = l < 30 ? C.plan.rowsites[(size_t)row_tile * RS + slab * 6 + l / 5] : -1;
            row_groups[r] = site >= 0 ? 5 * site + l % 5 : -1;
        }
    if (col_groups)
        for (int c = 0; c < NC; c++) {
            const int64_t g = (int64_t)col_tile * NC + c;
            col_groups[c] = g < (int64_t)5 * pk->N ? (int32_t)g : -1;
        }
    return RR_OK;
}

// measurement hook (include/rr_debug.h): bare tcgen05.mma loop of the scan's MMA kind on every SM, best of `reps`
extern "C" int rr_debug_mma_peak_shape(int device, int variant, int cta_group, int n_cols, int kblocks_per_sm, int reps, float *best_ms, double *macs);
extern "C" int rr_debug_mma_peak(int device, int variant, int kblocks_per_sm, int reps, float *best_ms, double *macs)
{
    return rr_debug_mma_peak_shape(device, variant, 1, 240, kblocks_per_sm, reps, best_ms, macs);
}

extern "C" int rr_debug_mma_peak_shape(int device, int variant, int cta_group, int n_cols, int kblocks_per_sm, int reps, float *best_ms, double *macs)
{
    if (!best_ms || !macs || reps < 1 || kblocks_per_sm < 1) { rr_set_error("rr_debug_mma_peak: bad arguments"); return RR_E_ARG; }
    if (!(cta_group == 1 && n_cols == 240) && !(cta_group == 2 && n_cols >= 32 && n_cols <= 256 && n_cols % 16 == 0)) {
        rr_set_error("rr_debug_mma_peak_shape: cta_group 1 with 240 columns, or cta_group 2 with 32..256 columns in steps of 16");
        return RR_E_ARG;
    }
    if (variant != RR_VARIANT_UMMA && variant != RR_VARIANT_UMMA_F4 && variant != RR_VARIANT_UMMA_MXF4) { rr_set_error("rr_debug_mma_peak: a tcgen05 variant is required"); return RR_E_ARG; }
    const int ndev = rr_device_count();
    if (ndev <= 0) { rr_set_error("no CUDA device"); return RR_E_NODEV; }
    if (device < 0 || device >= ndev) { rr_set_error("device %d out of range", device); return RR_E_ARG; }
    RR_CUDA(cudaSetDevice(device));
    int n_sm = 0;
    RR_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    const int mode = variant == RR_VARIANT_UMMA_MXF4 ? 2 : variant == RR_VARIANT_UMMA_F4 ? 1 : 0;
    event_scope ev;
    cudaEvent_t a, b;
    if (ev.create(&a) || ev.create(&b)) { rr_set_error("cudaEventCreate failed"); return RR_E_CUDA; }
    *best_ms = 0.f;
    for (int r = 0; r <= reps; r++) {                                    // r = 0: warm-up
        float ms = 0.f;
        RR_CUDA(cudaEventRecord(a, nullptr));
        if (cta_group == 2) RR_CUDA(rr_umma_mma_peak_pair(mode, n_sm, kblocks_per_sm, n_cols, macs, nullptr));
        else RR_CUDA(rr_umma_mma_peak(mode, n_sm, kblocks_per_sm, macs, nullptr));
        RR_CUDA(cudaEventRecord(b, nullptr));
        RR_CUDA(cudaEventSynchronize(b));
        RR_CUDA(cudaEventElapsedTime(&ms, a, b));
        if (r > 0 && (*best_ms == 0.f || ms < *best_ms)) *best_ms = ms;
    }
    return RR_OK;
}

// ---------------------------------------------------------------------------------------
// the whole path: pack on n GPUs, scan one part each, merge (882-891)
// ---------------------------------------------------------------------------------------
extern "C" int rr_maxcorr_run(const rr_msa *msa, int mincov, int n_gpus, int variant, unsigned flags,
                              double *maxcorr_out, int32_t *argmax_out, rr_scan_stats *stats)
{
    if (!msa || (!maxcorr_out && msa->cols > 0) || n_gpus < 1) { rr_set_error("rr_maxcorr_run: bad arguments"); return RR_E_ARG; }
    RR_TRACE("run: start");
    const int ndev = rr_device_count();
    RR_TRACE("run: device count");
    if (ndev <= 0) { rr_set_error("no CUDA device: the scan has no CPU fallback"); return RR_E_NODEV; }
    if (n_gpus > ndev) { rr_set_error("%d GPUs requested, %d present", n_gpus, ndev); return RR_E_ARG; }
    const size_t G = (size_t)5 * msa->cols;
    std::vector<std::vector<double>> M(n_gpus, std::vector<double>(G));
    std::vector<std::vector<int32_t>> A(n_gpus, std::vector<int32_t>(G));
    std::vector<rr_scan_stats> S(n_gpus);
    std::vector<int> RC(n_gpus, RR_OK);
    std::vector<std::string> ERR(n_gpus);
    std::vector<rr_packed *> PK(n_gpus, nullptr);

    // phase barrier for the per-GPU host threads
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0, generation = 0;
    auto barrier = [&]() {
        std::unique_lock<std::mutex> lk(mu);
        const int gen = generation;
        if (++arrived == n_gpus) { arrived = 0; generation++; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen; });
    };
    // spans of all rows, filled slice by slice by the workers; per-GPU export buffers of the seeded maxima
    std::vector<int32_t> spans((size_t)3 * std::max(msa->rows, 1), 0);
    std::vector<double *> d_vals(n_gpus, nullptr);
    auto all_ok = [&]() { for (int e = 0; e < n_gpus; e++) if (RC[e] != RR_OK) return false; return true; };

    auto worker = [&](int d) {
        rr_scan_opts o;
        const unsigned scan_flags = flags & ~RR_FLAG_HOST_FINALIZE;
        o.mincov = mincov; o.variant = variant; o.flags = scan_flags; o.part_index = d; o.part_count = n_gpus;
        int rc;
        if (n_gpus == 1) {
            rc = rr_pack(msa, d, &PK[d]);
        } else {
            // ---- the packing is shared: GPU d uploads and packs the rows [lo, hi) only, the bitsets are merged over NVLink
            // (reduce-scatter + all-gather with peer copies and a word-wise OR), then every GPU finishes on the whole MSA
            const int R = msa->rows, lo = (int)((int64_t)R * d / n_gpus), hi = (int)((int64_t)R * (d + 1) / n_gpus);
            rc = rr_pack_rows(msa, d, lo, hi, &PK[d]);
            if (!rc) rc = rr_pack_slice_spans(PK[d], spans.data() + lo, spans.data() + R + lo, spans.data() + 2 * (size_t)R + lo);
            RC[d] = rc;
            barrier();
            if (!all_ok()) { if (rc) ERR[d] = rr_last_error(); return; }
            rc = rr_pack_set_spans(PK[d], spans.data(), spans.data() + R, spans.data() + 2 * (size_t)R);
            RC[d] = rc;
            barrier();                                                   // every slice is packed (the call drains its stream)
            if (!all_ok()) { if (rc) ERR[d] = rr_last_error(); return; }
            rr_packed *me = PK[d];
            cudaSetDevice(d);
            rr_alloc_stream(me->st);
            for (int e = 0; e < n_gpus; e++)
                if (e != d && cudaDeviceEnablePeerAccess(e, 0) != cudaSuccess) cudaGetLastError();   // already enabled / not possible: copies still work
            const size_t total16 = (size_t)6 * me->N * me->W32 * sizeof(uint32_t) / 16;   // W32 is a multiple of 4
            auto chunk_lo = [&](int c) { return total16 * (size_t)c / n_gpus * 16; };
            const size_t my_lo = chunk_lo(d), my_bytes = chunk_lo(d + 1) - my_lo;
            cudaError_t ce = cudaSuccess;
            {
                dev_scope scope;
                uint8_t *tmp = nullptr;
                rc = scope.alloc(&tmp, std::max<size_t>(my_bytes, 16));
                for (int k = 1; k < n_gpus && !rc && ce == cudaSuccess; k++) {   // my chunk of every peer's buffer, ORed into mine
                    const int e = (d + k) % n_gpus;
                    ce = cudaMemcpyPeerAsync(tmp, d, (const uint8_t *)PK[e]->d_bits + my_lo, e, my_bytes, me->st);
                    if (ce == cudaSuccess) ce = rr_launch_or_words((uint8_t *)me->d_bits + my_lo, tmp, my_bytes, me->st);
                }
                if (ce == cudaSuccess) ce = cudaStreamSynchronize(me->st);
            }
            if (ce != cudaSuccess && !rc) { rr_set_error("CUDA error %s while merging the packed bitsets", cudaGetErrorString(ce)); rc = RR_E_CUDA; }
            RC[d] = rc;
            barrier();                                                   // every GPU's own chunk is final
            if (!all_ok()) { if (rc) ERR[d] = rr_last_error(); return; }
            for (int k = 1; k < n_gpus && ce == cudaSuccess; k++) {      // the other chunks from their owners
                const int e = (d + k) % n_gpus;
                const size_t lo_e = chunk_lo(e), bytes_e = chunk_lo(e + 1) - lo_e;
                ce = cudaMemcpyPeerAsync((uint8_t *)me->d_bits + lo_e, d, (const uint8_t *)PK[e]->d_bits + lo_e, e, bytes_e, me->st);
            }
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(me->st);
            if (ce != cudaSuccess) { rr_set_error("CUDA error %s while gathering the packed bitsets", cudaGetErrorString(ce)); rc = RR_E_CUDA; }
            RC[d] = rc;
            barrier();                                                   // nobody reads a peer's buffer any more
            if (!all_ok()) { if (rc) ERR[d] = rr_last_error(); return; }
            rc = rr_pack_finish(PK[d]);
            // ---- seeding pass on every GPU (column chunks split between them); the maxima are exchanged device to device and
            // raised on every GPU as common thresholds (values only: they prune, and lose every tie against a real pair)
            if (!rc) { o.flags = scan_flags | RR_FLAG_SEED_ONLY; rc = rr_scan(PK[d], &o, &S[d]); }
            const size_t Gd = (size_t)5 * msa->cols;
            if (!rc) rc = dev_alloc(&d_vals[d], Gd);
            if (!rc && rr_launch_best_values(me->d_best, d_vals[d], (int64_t)Gd, me->st) != cudaSuccess) { rr_set_error("CUDA error exporting the seeded maxima"); rc = RR_E_CUDA; }
            if (!rc && cudaStreamSynchronize(me->st) != cudaSuccess) { rr_set_error("CUDA error exporting the seeded maxima"); rc = RR_E_CUDA; }
            RC[d] = rc;
            barrier();
            if (!all_ok()) { if (rc) ERR[d] = rr_last_error(); rr_dev_free(d_vals[d]); return; }
            {
                dev_scope scope;
                double *tmp = nullptr;
                rc = scope.alloc(&tmp, Gd);
                for (int k = 1; k < n_gpus && !rc && ce == cudaSuccess; k++) {
                    const int e = (d + k) % n_gpus;
                    ce = cudaMemcpyPeerAsync(tmp, d, d_vals[e], e, Gd * sizeof(double), me->st);
                    if (ce == cudaSuccess) ce = rr_launch_raise_best(me->d_best, tmp, (int64_t)Gd, me->st);
                }
                if (ce == cudaSuccess) ce = cudaStreamSynchronize(me->st);
            }
            if (ce != cudaSuccess && !rc) { rr_set_error("CUDA error %s while exchanging the seeded maxima", cudaGetErrorString(ce)); rc = RR_E_CUDA; }
            RC[d] = rc;
            barrier();                                                   // peers are done reading d_vals[d]
            rr_dev_free(d_vals[d]);
            if (!all_ok()) { if (rc) ERR[d] = rr_last_error(); return; }
            o.flags = scan_flags | RR_FLAG_SKIP_SEED;
        }
        if (!rc) rc = rr_scan(PK[d], &o, &S[d]);
        if (!rc) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a, PK[d]->st);
            rc = rr_scan_fetch(PK[d], M[d].data(), A[d].data());
            cudaEventRecord(b, PK[d]->st);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&S[d].fetch_ms, a, b);
            cudaEventDestroy(a); cudaEventDestroy(b);
        }
        RC[d] = rc;
        if (rc) ERR[d] = rr_last_error();
    };
    if (n_gpus == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < n_gpus; d++) th.emplace_back(worker, d);
        for (auto &t : th) t.join();
    }
    RR_TRACE("run: scans fetched");
    int rc = RR_OK;
    for (int d = 0; d < n_gpus; d++)
        if (RC[d]) { rc = RC[d]; rr_set_error("GPU %d: %s", d, ERR[d].c_str()); break; }

    if (!rc) {
        for (size_t g = 0; g < G; g++) {
            double z = M[0][g];
            int32_t p = A[0][g];
            for (int d = 1; d < n_gpus; d++) {
                if (M[d][g] > z) { z = M[d][g]; p = A[d][g]; }
                else if (M[d][g] == z && z > 0.0 && A[d][g] >= 0 && (p < 0 || A[d][g] < p)) p = A[d][g];
            }
            maxcorr_out[g] = z;
            if (argmax_out) argmax_out[g] = p;
            A[0][g] = p;
        }
        if (stats) {
            *stats = S[0];
            for (int d = 1; d < n_gpus; d++) {
                stats->pair_tests += S[d].pair_tests; stats->exact_evals += S[d].exact_evals;
                stats->bound_evals += S[d].bound_evals; stats->work_units += S[d].work_units;
                stats->executed_ops += S[d].executed_ops;
                stats->kernel_ms = std::max(stats->kernel_ms, S[d].kernel_ms);
                stats->h2d_ms = std::max(stats->h2d_ms, S[d].h2d_ms);
                stats->pack_ms = std::max(stats->pack_ms, S[d].pack_ms);
                stats->prepare_ms = std::max(stats->prepare_ms, S[d].prepare_ms);
                stats->fetch_ms = std::max(stats->fetch_ms, S[d].fetch_ms);
            }
        }
        RR_TRACE("run: merged");
        if (flags & RR_FLAG_HOST_FINALIZE) {
            cudaEvent_t a, b;
            cudaSetDevice(PK[0]->device);
            cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a, PK[0]->st);
            rc = rr_scan_finalize(PK[0], maxcorr_out, A[0].data());
            cudaEventRecord(b, PK[0]->st);
            cudaEventSynchronize(b);
            if (stats) cudaEventElapsedTime(&stats->finalize_ms, a, b);
            cudaEventDestroy(a); cudaEventDestroy(b);
        }
    }
    RR_TRACE("run: finalized");
    for (int d = 0; d < n_gpus; d++) rr_packed_free(PK[d]);
    RR_TRACE("run: freed");
    return rc;
}
