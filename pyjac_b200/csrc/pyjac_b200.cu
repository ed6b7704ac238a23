// C ABI of the B200-native evaluator (include/pyjac_b200.h): mechanism handle, launch
// logic, the host-pointer batch API that stands in for pyjac/pywrap/pyjacob.cu:84-188 and
// the reference-named scalar entry points.  There is no CPU compute path in this file.
#include "pyjac_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "eval.cuh"
#include "jac6.cuh"
#include "consumer.cuh"
#include "pjtable.h"

#include <map>

using pj::IO;
using pj::Tables;

struct pyjac_mech {
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    int smem_per_sm = 0;
    Tables tb{};
    pj5::Plan plan{};
    pj6::Plan6 plan6{};              // record streams of k_jac6 (eval_jacob of plans that live in shared memory)
    bool has6 = false;
    std::vector<void*> dev_allocs;
    int user_bpsm = 0;
    int bpsm[4] = {0, 0, 0, 0};      // blocks per SM per mode (0 = not configured yet)
    // factored output (M_FACT): pattern of the sparse block in record order (host copies) and by row (device)
    std::vector<int> fac_rows, fac_cols;
    std::vector<double> colfac_h;
    pjc::Fac fac{};
    std::atomic<long long> launches{0};
    std::mutex mu;                   // launch configuration and the shared scratch of a wsg plan
    std::vector<int> fwd_map, back_map;   // species: internal position -> original index and back (apply_mask)
    int conv = 0;                    // 1: the reference-named dydt is the constant-volume one (header.h: #define CONV)
    char* ws = nullptr;              // per-block working sets of a plan with wsg = 1
    size_t ws_bytes = 0;
    cudaEvent_t ws_ev = nullptr;     // orders the launches that share `ws` across streams
    // staging for the host-pointer API
    cudaStream_t stream[2] = {nullptr, nullptr};
    double* h_pin[2] = {nullptr, nullptr};
    double* d_in[2] = {nullptr, nullptr};
    double* d_out[2] = {nullptr, nullptr};
    size_t pin_bytes = 0, din_bytes = 0, dout_bytes = 0;
};

namespace {

thread_local std::string g_err;
std::mutex g_mu;
pyjac_mech* g_current = nullptr;    // mechanism behind surfaces 2 and 3
int g_cu_num = 0;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? PYJAC_ENOMEM : PYJAC_ECUDA,     \
                        std::string(#call) + ": " + cudaGetErrorString(e_));             \
    } while (0)

template <int MAXT, int MODE>
const void* kernel_t(int gs)
{
    switch (gs) {
#ifndef PJ_DEV_GS8_ONLY          // development builds (tools/): only the GRI-sized instantiation
    case 2: return (const void*)pj5::k_eval<2, MAXT, MODE>;
    case 4: return (const void*)pj5::k_eval<4, MAXT, MODE>;
    case 16: return (const void*)pj5::k_eval<16, MAXT, MODE>;
    case 32: return (const void*)pj5::k_eval<32, MAXT, MODE>;
#endif
    case 8: return (const void*)pj5::k_eval<8, MAXT, MODE>;
    default: return nullptr;
    }
}

template <int MAXT, int MODE>
const void* kernel_g(int gs)
{
    switch (gs) {
#ifndef PJ_DEV_GS8_ONLY
    case 2: return (const void*)pj5::k_eval<2, MAXT, MODE, true>;
    case 16: return (const void*)pj5::k_eval<16, MAXT, MODE, true>;
    case 32: return (const void*)pj5::k_eval<32, MAXT, MODE, true>;
#endif
    case 8: return (const void*)pj5::k_eval<8, MAXT, MODE, true>;
    default: return nullptr;
    }
}

// k_eval for a plan (states per block, block size <= 512) and a mode.  Blocks of up to 384
// threads get the 168-register build, larger ones the 128-register build.  Plans whose working
// set lives in global memory (wsg) exist for 2, 8, 16 and 32 states per block and up to 384 threads.
const void* kernel_for(int gs, int mode, int nt, int wsg = 0)
{
    if (wsg) {
        // no shared memory in use: blocks of up to 256 threads take the 128-register build so that two
        // of them are resident per SM (their phases interleave), larger ones the 168-register build
        if (nt > 384) return nullptr;
        if (nt <= 256)
            return mode == pj::M_DYDT ? kernel_g<512, pj::M_DYDT>(gs) : mode == pj::M_RATES ? kernel_g<512, pj::M_RATES>(gs)
                 : mode == pj::M_FACT ? kernel_g<512, pj::M_FACT>(gs) : kernel_g<512, pj::M_JAC>(gs);
        return mode == pj::M_DYDT ? kernel_g<384, pj::M_DYDT>(gs) : mode == pj::M_RATES ? kernel_g<384, pj::M_RATES>(gs)
             : mode == pj::M_FACT ? kernel_g<384, pj::M_FACT>(gs) : kernel_g<384, pj::M_JAC>(gs);
    }
    const bool small = nt <= 384;
    if (mode == pj::M_DYDT) return small ? kernel_t<384, pj::M_DYDT>(gs) : kernel_t<512, pj::M_DYDT>(gs);
    if (mode == pj::M_RATES) return small ? kernel_t<384, pj::M_RATES>(gs) : kernel_t<512, pj::M_RATES>(gs);
    if (mode == pj::M_FACT) return small ? kernel_t<384, pj::M_FACT>(gs) : kernel_t<512, pj::M_FACT>(gs);
    return small ? kernel_t<384, pj::M_JAC>(gs) : kernel_t<512, pj::M_JAC>(gs);
}

// eval_jacob of a plan with record streams (p6_*): k_jac6, 168-register build up to 384 threads
const void* kernel6_for(int gs, int nt)
{
    const bool small = nt <= 384;
    switch (gs) {
#ifndef PJ_DEV_GS8_ONLY
    case 4: return small ? (const void*)pj6::k_jac6<4, 384> : (const void*)pj6::k_jac6<4, 512>;
    case 16: return small ? (const void*)pj6::k_jac6<16, 384> : (const void*)pj6::k_jac6<16, 512>;
    case 32: return small ? (const void*)pj6::k_jac6<32, 384> : (const void*)pj6::k_jac6<32, 512>;
#endif
    case 8: return small ? (const void*)pj6::k_jac6<8, 384> : (const void*)pj6::k_jac6<8, 512>;
    default: return nullptr;
    }
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel instantiation, which handles with
// different plans share: keep it at the largest size any handle of this process asked for
int ensure_dyn_smem(const void* fn, int device, size_t bytes, bool prefer_l1)
{
    static std::map<std::pair<const void*, int>, size_t> granted;       // function attributes are per device
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = granted.find({fn, device});
    if (it == granted.end()) {
        CU(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                prefer_l1 ? cudaSharedmemCarveoutMaxL1 : cudaSharedmemCarveoutMaxShared));
        it = granted.emplace(std::make_pair(fn, device), (size_t)0).first;
    }
    if (bytes > it->second) {
        CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        it->second = bytes;
    }
    return PYJAC_OK;
}

// restores the caller's current device when an entry point returns
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// One launch of k_eval; the plan in the table blob fixes states per block and block size.
int launch(pyjac_mech* m, int mode, const IO& io_in, cudaStream_t st)
{
    if (io_in.n <= 0) return PYJAC_OK;
    IO io = io_in;
    DeviceGuard guard(m->device);
    const pj5::Plan& pl = m->plan;
    if (mode == pj::M_JAC && m->has6) {
        // the Jacobian of a plan whose working set lives in shared memory: record-stream kernel
        const pj6::Plan6& p6 = m->plan6;
        const void* fn6 = kernel6_for(p6.gs, p6.nt);
        if (!fn6) return fail(PYJAC_EINVAL, "table blob holds no usable stream plan");
        if ((size_t)p6.bytes > (size_t)m->smem_optin)
            return fail(PYJAC_ETOOBIG, "mechanism working set does not fit in shared memory");
        int rc6 = ensure_dyn_smem(fn6, m->device, (size_t)p6.bytes, false);
        if (rc6) return rc6;
        const long long groups6 = ((long long)io.n + p6.gs - 1) / p6.gs;
        int grid6 = (int)std::min<long long>(groups6, (long long)m->sm_count);
#ifdef PJ_DEV
        if (m->user_bpsm < 0) grid6 = std::min(grid6, -m->user_bpsm);      // development: explicit grid
#endif
        void* args6[3] = {(void*)&m->tb, (void*)&m->plan6, (void*)&io};
        CU(cudaLaunchKernel(fn6, dim3(grid6), dim3(p6.nt), args6, (size_t)p6.bytes, st));
        ++m->launches;
        return PYJAC_OK;
    }
    const void* fn = kernel_for(pl.gs, mode, pl.nt, pl.wsg);
    if (!fn) return fail(PYJAC_EINVAL, "table blob holds no usable plan");
    const size_t bytes = pl.wsg ? 0 : (size_t)pl.total * 8;
    if (bytes > (size_t)m->smem_optin)
        return fail(PYJAC_ETOOBIG, "mechanism working set does not fit in shared memory");
    {
        int rcs = ensure_dyn_smem(fn, m->device, bytes, pl.wsg != 0);
        if (rcs) return rcs;
    }
    std::unique_lock<std::mutex> lk(m->mu);          // held to the end for a wsg plan (shared scratch), else only here
    if (!m->bpsm[mode]) {
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, pl.nt, bytes));
        if (occ < 1) return fail(PYJAC_ETOOBIG, "kernel cannot be resident with this plan");
        if (m->user_bpsm) occ = std::min(occ, m->user_bpsm);
        m->bpsm[mode] = occ;
    }
    const long long groups = ((long long)io.n + pl.gs - 1) / pl.gs;
    const int grid = (int)std::min<long long>(groups, (long long)m->sm_count * m->bpsm[mode]);
    if (!pl.wsg) lk.unlock();
    if (pl.wsg) {
        // one working set per resident block, in global memory (L1 / L2 hold what is hot); the
        // launches of one handle share it, so an event orders them across streams
        const size_t need = (size_t)m->sm_count * m->bpsm[mode] * (size_t)pl.total * 8;
        if (need > m->ws_bytes) {
            if (m->ws) { CU(cudaDeviceSynchronize()); cudaFree(m->ws); m->ws = nullptr; m->ws_bytes = 0; }
            CU(cudaMalloc(&m->ws, need));
            m->ws_bytes = need;
        }
        io.ws = m->ws;
        if (!m->ws_ev) CU(cudaEventCreateWithFlags(&m->ws_ev, cudaEventDisableTiming));
        else CU(cudaStreamWaitEvent(st, m->ws_ev, 0));
    }
    void* args[3] = {(void*)&m->tb, (void*)&m->plan, (void*)&io};
    CU(cudaLaunchKernel(fn, dim3(grid), dim3(pl.nt), args, bytes, st));
    if (pl.wsg) CU(cudaEventRecord(m->ws_ev, st));
    ++m->launches;
    return PYJAC_OK;
}


// Pattern of the factored record's sparse block (plan.py: fac_rows / fac_cols, column-major record order):
// checked, kept on the host for pyjac_factored_pattern and uploaded by row for the consumers.
int setup_factored(pyjac_mech* m, const void* blob)
{
    const int nsp = m->tb.nsp;
    const pjt::Entry* pe = pjt::find(blob, "p5_cfg");
    const pjt::Entry* re = pjt::find(blob, "fac_rows");
    const pjt::Entry* ce = pjt::find(blob, "fac_cols");
    const pjt::Entry* fe = pjt::find(blob, "p5_colfac");
    if (!pe || pe->count < 16 || !re || !ce || re->dtype != 1 || ce->dtype != 1 || !fe)
        return fail(PYJAC_EINVAL, "table blob lacks the factored-output pattern");
    const int nnz = ((const int*)((const char*)blob + pe->offset))[15];
    if (nnz < 0 || re->count < nnz || ce->count < nnz || (long long)nnz > (long long)nsp * nsp)
        return fail(PYJAC_EINVAL, "bad factored-output pattern");
    const int* rows = (const int*)((const char*)blob + re->offset);
    const int* cols = (const int*)((const char*)blob + ce->offset);
    m->fac_rows.assign(rows, rows + nnz);
    m->fac_cols.assign(cols, cols + nnz);
    const double* cf = (const double*)((const char*)blob + fe->offset);
    m->colfac_h.assign(cf, cf + 2 * (size_t)nsp);
    std::vector<int> ptr(nsp, 0), slot(std::max(nnz, 1), 0), col(std::max(nnz, 1), 0);
    for (int p = 0; p < nnz; ++p) {
        if (rows[p] < 1 || rows[p] >= nsp || cols[p] < 1 || cols[p] >= nsp) return fail(PYJAC_EINVAL, "bad factored-output pattern");
        ++ptr[rows[p]];                      // ptr[k + 1] counts row k + 1 (k = 0 .. nsp-2)
    }
    for (int k = 1; k < nsp; ++k) ptr[k] += ptr[k - 1];
    std::vector<int> fill(ptr.begin(), ptr.end());      // row k + 1 starts at ptr[k]
    for (int p = 0; p < nnz; ++p) {
        const int at = fill[rows[p] - 1]++;
        slot[at] = nsp + 3 * (nsp - 1) + p;
        col[at] = cols[p];
    }
    auto up = [&](const std::vector<int>& v, const int** out) -> int {
        void* dptr = nullptr;
        CU(cudaMalloc(&dptr, std::max<size_t>(v.size() * 4, 64)));
        m->dev_allocs.push_back(dptr);
        CU(cudaMemcpy(dptr, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
        *out = (const int*)dptr;
        return PYJAC_OK;
    };
    pjc::Fac& f = m->fac;
    f.nsp = nsp; f.nnz = nnz; f.colfac = m->plan.colfac;
    std::vector<int> r_(m->fac_rows), c_(m->fac_cols);
    if (r_.empty()) { r_.push_back(0); c_.push_back(0); }
    int rc = up(ptr, &f.ptr);
    if (!rc) rc = up(slot, &f.slot);
    if (!rc) rc = up(col, &f.col);
    if (!rc) rc = up(r_, &f.rows);
    if (!rc) rc = up(c_, &f.cols);
    m->plan.fac_nnz = nnz;
    return rc;
}

template <typename T>
int upload(pyjac_mech* m, const void* blob, const char* name, const T** out, int dtype)
{
    const pjt::Entry* e = pjt::find(blob, name);
    if (!e || e->dtype != dtype) return fail(PYJAC_EINVAL, std::string("table blob lacks ") + name);
    const size_t payload = (size_t)e->count * pjt::elem_size(e->dtype);
    const size_t bytes = std::max<size_t>((payload + 15) / 16 * 16, 64);
    void* d = nullptr;
    CU(cudaMalloc(&d, bytes));
    m->dev_allocs.push_back(d);
    CU(cudaMemset(d, 0, bytes));
    CU(cudaMemcpy(d, (const char*)blob + e->offset, payload, cudaMemcpyHostToDevice));
    *out = (const T*)d;
    return PYJAC_OK;
}

void release_staging(pyjac_mech* m)
{
    for (int i = 0; i < 2; ++i) {
        if (m->h_pin[i]) cudaFreeHost(m->h_pin[i]);
        if (m->d_in[i]) cudaFree(m->d_in[i]);
        if (m->d_out[i]) cudaFree(m->d_out[i]);
        m->h_pin[i] = m->d_in[i] = m->d_out[i] = nullptr;
    }
    m->pin_bytes = m->din_bytes = m->dout_bytes = 0;
}

int ensure_staging(pyjac_mech* m, size_t pin, size_t din, size_t dout)
{
    DeviceGuard guard(m->device);
    for (int i = 0; i < 2; ++i)
        if (!m->stream[i]) CU(cudaStreamCreateWithFlags(&m->stream[i], cudaStreamNonBlocking));
    if (pin > m->pin_bytes || din > m->din_bytes || dout > m->dout_bytes) {
        for (int i = 0; i < 2; ++i) CU(cudaStreamSynchronize(m->stream[i]));
        pin = std::max(pin, m->pin_bytes); din = std::max(din, m->din_bytes); dout = std::max(dout, m->dout_bytes);
        release_staging(m);
        for (int i = 0; i < 2; ++i) {
            CU(cudaMallocHost((void**)&m->h_pin[i], pin));
            CU(cudaMalloc((void**)&m->d_in[i], din));
            CU(cudaMalloc((void**)&m->d_out[i], dout));
        }
        m->pin_bytes = pin; m->din_bytes = din; m->dout_bytes = dout;
    }
    return PYJAC_OK;
}

// ---- staging of the scalar entry points: one slot per concurrent caller.  The reference's eval_jacob &c.
// are re-entrant (tester.c.in:24-29 calls eval_jacob inside an OpenMP loop); here every call borrows a
// slot (stream + pinned and device buffers) from a pool, so concurrent host threads do not share staging.
struct Slot {
    int device = -1;
    cudaStream_t st = nullptr;
    double *h = nullptr, *d_in = nullptr, *d_out = nullptr;
    size_t hb = 0, ib = 0, ob = 0;
};
std::vector<Slot*> g_slots;        // free slots
std::mutex g_slot_mu;

void slot_free(Slot* s)
{
    if (s->h) cudaFreeHost(s->h);
    if (s->d_in) cudaFree(s->d_in);
    if (s->d_out) cudaFree(s->d_out);
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
}

int slot_acquire(int device, size_t hb, size_t ib, size_t ob, Slot** out)
{
    Slot* s = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_slot_mu);
        for (size_t i = 0; i < g_slots.size(); ++i)
            if (g_slots[i]->device == device) { s = g_slots[i]; g_slots.erase(g_slots.begin() + i); break; }
    }
    if (!s) { s = new Slot(); s->device = device; }
    *out = s;
    if (!s->st) CU(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
    if (hb > s->hb) { if (s->h) cudaFreeHost(s->h); s->h = nullptr; s->hb = 0; CU(cudaMallocHost((void**)&s->h, hb)); s->hb = hb; }
    if (ib > s->ib) { if (s->d_in) cudaFree(s->d_in); s->d_in = nullptr; s->ib = 0; CU(cudaMalloc((void**)&s->d_in, ib)); s->ib = ib; }
    if (ob > s->ob) { if (s->d_out) cudaFree(s->d_out); s->d_out = nullptr; s->ob = 0; CU(cudaMalloc((void**)&s->d_out, ob)); s->ob = ob; }
    return PYJAC_OK;
}

void slot_release(Slot* s)
{
    std::lock_guard<std::mutex> lk(g_slot_mu);
    g_slots.push_back(s);
}

struct SlotLease {                 // returns the slot to the pool when the entry point leaves
    Slot* s = nullptr;
    ~SlotLease() { if (s) slot_release(s); }
};

// tables registered by a per-mechanism stub library (libgen.generate_library): the mechanism the
// reference-named entry points use when none was selected explicitly
const void* g_reg_blob = nullptr;
size_t g_reg_len = 0;

}  // namespace

extern "C" {

const char* pyjac_last_error(void) { return g_err.c_str(); }

int pyjac_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pyjac_mech_create(const void* blob, size_t len, int device, pyjac_mech** out)
{
    if (!out) return fail(PYJAC_EINVAL, "out is NULL");
    *out = nullptr;
    if (!pjt::valid(blob, len)) return fail(PYJAC_EINVAL, "not a PJB200T1 table blob");
    const pjt::Entry* de = pjt::find(blob, "dims");
    const pjt::Entry* ce = pjt::find(blob, "cst");
    if (!de || de->dtype != 1 || de->count < 16 || !ce || ce->dtype != 0 || ce->count < 2)
        return fail(PYJAC_EINVAL, "table blob lacks dims / cst");
    const int* d = (const int*)((const char*)blob + de->offset);
    const double* c = (const double*)((const char*)blob + ce->offset);
    {
        // a blob written by another version of tables.py / plan.py would be misread silently: refuse it
        const pjt::Entry* me = pjt::find(blob, "meta");
        const int* mv = me && me->dtype == 1 && me->count >= 3 ? (const int*)((const char*)blob + me->offset) : nullptr;
        if (!mv || mv[0] != pjt::SCHEMA_VERSION || mv[1] != pjt::PLAN5_VERSION ||
            (pjt::find(blob, "p6_cfg") && mv[2] != pjt::PLAN6_VERSION))
            return fail(PYJAC_EINVAL, "table blob was written for another version of the library (rebuild the tables)");
        // table lengths against the dimensions they are indexed with
        struct { const char* name; long long count; } want[] = {
            {"sp_w", d[0]}, {"sp_iw", d[0]}, {"sp_ruw", d[0]}, {"sp_tmid", d[0]}, {"sp_nasa", 32LL * d[0]},
            {"p5_rx", 16LL * d[1]}, {"p5_rxout", 4LL * d[1]}, {"red_off", d[0] + 1LL}, {"plog_off", d[1] + 1LL},
            {"cheb_off", d[1] + 1LL}, {"sp_fwd_map", d[0]}, {"p5_colfac", 2LL * d[0]},
            {"p5_fac_map", (long long)d[0] * d[0]}};
        for (const auto& w_ : want) {
            const pjt::Entry* e = pjt::find(blob, w_.name);
            if (!e || e->count < w_.count) return fail(PYJAC_EINVAL, std::string("table blob: ") + w_.name + " is missing or too short");
        }
        if (d[0] < 2 || d[1] < 1 || d[4] < 0 || d[8] < 0 || d[8] > d[1] || d[9] != d[1] - d[8])
            return fail(PYJAC_EINVAL, "table blob: inconsistent dimensions");
    }
    if (pyjac_device_count() <= 0) return fail(PYJAC_ENODEVICE, "no CUDA device available (no CPU fallback exists)");
    if (device < 0) CU(cudaGetDevice(&device));
    DeviceGuard guard(device);
    pyjac_mech* m = new pyjac_mech();
    m->device = device;
    {
        const pjt::Entry* me = pjt::find(blob, "meta");
        m->conv = me->count > 3 ? ((const int*)((const char*)blob + me->offset))[3] != 0 : 0;
    }
    Tables& t = m->tb;
    t.nsp = d[0]; t.nr = d[1]; t.nrev = d[2]; t.npd = d[3]; t.nraw = d[4];
    t.first_pm = d[8]; t.npm = d[9]; t.nplog = d[5]; t.ncheb = d[6];
    t.ru = c[0];
    int rc = PYJAC_OK;
#define UP(field, name, type, code) if (!rc) rc = upload<type>(m, blob, name, &m->field, code)
    UP(tb.sp_w, "sp_w", double, 0); UP(tb.sp_iw, "sp_iw", double, 0); UP(tb.sp_ruw, "sp_ruw", double, 0);
    UP(tb.sp_tmid, "sp_tmid", double, 0); UP(tb.sp_nasa, "sp_nasa", double, 0);
    UP(tb.pm_par, "pm_par", double, 0); UP(tb.pm_sp, "pm_sp", int, 1);
    UP(tb.red_off, "red_off", int, 1); UP(tb.red_rx, "red_rx", int, 1); UP(tb.red_nu, "red_nu", double, 0);
    UP(tb.rx_out, "p5_rxout", int4, 1);
    UP(tb.plog_off, "plog_off", int, 1); UP(tb.plog_par, "plog_par", double, 0);
    UP(tb.cheb_off, "cheb_off", int, 1); UP(tb.cheb_par, "cheb_par", double, 0);
    if (!rc) {
        const pjt::Entry* pe = pjt::find(blob, "p5_cfg");
        if (!pe || pe->dtype != 1 || pe->count < 14) rc = fail(PYJAC_EINVAL, "table blob lacks p5_cfg");
        else {
            const int* c5 = (const int*)((const char*)blob + pe->offset);
            pj5::Plan& pl = m->plan;
            int* o = &pl.gs;
            for (int i = 0; i < 14; ++i) o[i] = c5[i];
            pl.wsg = pe->count > 14 ? c5[14] : 0;
            // the regions of the working set (rows of gs doubles) must lie inside `total`
            const long long gs_ = pl.gs, nraw_ = d[4];
            const bool regions_ok = pl.oSP >= 0 && pl.oSP + (d[0] + 1LL) * pj5::SP_SLOTS * gs_ <= pl.total &&
                                    pl.oRX >= 0 && pl.oRX + (d[1] + 2LL) * pj5::RX_SLOTS * gs_ <= pl.total &&
                                    pl.oRAW >= 0 && pl.oRAW + (nraw_ + 2) * gs_ <= pl.total &&
                                    pl.oSC >= 0 && pl.oSC + (long long)pj5::NQ * gs_ <= pl.total &&
                                    pl.oPA >= 0 && pl.oPA + (long long)pl.nw * pj5::NPART * gs_ <= pl.total &&
                                    pl.oCF >= 0 && pl.oCF + 2LL * d[0] <= pl.total;
            if (!kernel_for(pl.gs, pj::M_JAC, pl.nt, pl.wsg) || pl.nt != pl.nw * 32 || pl.nt < 64 || pl.nt > 512 || pl.nsub * pl.gs != 64 || pl.coop < 1 || pl.coop > pl.nsub || pl.tcoop < 1 || pl.tcoop > pl.nsub || !regions_ok)
                rc = fail(PYJAC_EINVAL, "bad plan configuration");
        }
    }
    if (!rc) m->plan.rx_out = m->tb.rx_out;
    if (!rc) {
        // species permutation of the mechanism (utils.get_species_mappings, utils.py:55-91) for apply_mask
        const pjt::Entry* fe = pjt::find(blob, "sp_fwd_map");
        if (!fe || fe->dtype != 1 || fe->count != t.nsp) rc = fail(PYJAC_EINVAL, "table blob lacks sp_fwd_map");
        else {
            const int* f = (const int*)((const char*)blob + fe->offset);
            m->fwd_map.assign(f, f + t.nsp);
            m->back_map.assign(t.nsp, -1);
            for (int i = 0; i < t.nsp && !rc; ++i) {
                if (f[i] < 0 || f[i] >= t.nsp || m->back_map[f[i]] >= 0) rc = fail(PYJAC_EINVAL, "sp_fwd_map is not a permutation");
                else m->back_map[f[i]] = i;
            }
        }
    }
    UP(plan.rx, "p5_rx", int4, 1); UP(plan.eff_off, "p5_eff_off", int, 1); UP(plan.eff, "p5_eff", int4, 1);
    UP(plan.b_off, "p5_b_off", int, 1); UP(plan.b_npm, "p5_b_npm", int, 1); UP(plan.b_item, "p5_b_item", int, 1);
    UP(plan.c_off, "p5_c_off", int, 1); UP(plan.c_item, "p5_c_item", int4, 1); UP(plan.c_str, "p5_c_str", uint2, 1);
    UP(plan.d_off, "p5_d_off", int, 1); UP(plan.d_item, "p5_d_item", int2, 1); UP(plan.d_str, "p5_d_str", uint2, 1);
    UP(plan.s_off, "p5_s_off", int, 1); UP(plan.s_str, "p5_s_str", uint4, 1);
    UP(plan.o_off, "p5_o_off", int, 1); UP(plan.o_str, "p5_o_str", uint2, 1);
    UP(plan.t_off, "p5_t_off", int, 1); UP(plan.t_item, "p5_t_item", int2, 1); UP(plan.t_str, "p5_t_str", uint2, 1);
    UP(plan.colfac, "p5_colfac", double2, 0);
    UP(plan.fac_map, "p5_fac_map", int, 1);
    if (!rc) rc = setup_factored(m, blob);
    if (!rc && pjt::find(blob, "p6_cfg")) {
        const pjt::Entry* pe = pjt::find(blob, "p6_cfg");
        if (pe->dtype != 1 || pe->count < 24) rc = fail(PYJAC_EINVAL, "bad p6_cfg");
        else {
            const int* c6 = (const int*)((const char*)blob + pe->offset);
            pj6::Plan6& p6 = m->plan6;
            int* o = &p6.gs;
            for (int i = 0; i < 24; ++i) o[i] = c6[i];
            if (!kernel6_for(p6.gs, p6.nt) || p6.nt != p6.nw * 32 || p6.nsub * p6.gs != 64 || p6.chb != pj6::CHB ||
                p6.nslot != pj6::NSLOT || p6.chr != pj6::CHB / (p6.nsub * 16) || p6.coop < 1 || p6.coop > p6.nsub ||
                p6.tcoop < 1 || p6.tcoop > p6.nsub || p6.bytes < p6.mbar + p6.nw * pj6::NSLOT * 8)
                rc = fail(PYJAC_EINVAL, "bad stream plan configuration");
        }
        UP(plan6.str, "p6_str", uint4, 1); UP(plan6.hdr, "p6_hdr", int, 1);
        UP(plan6.eff, "p6_eff", int4, 1); UP(plan6.colfac, "p6_colfac", double2, 0);
        if (!rc) { m->plan6.eff_off = m->plan.eff_off; m->has6 = true; }
    }
#undef UP
    if (!rc) {
        cudaDeviceProp prop;
        cudaError_t e = cudaGetDeviceProperties(&prop, device);
        if (e != cudaSuccess) rc = fail(PYJAC_ECUDA, cudaGetErrorString(e));
        else { m->sm_count = prop.multiProcessorCount; m->smem_optin = (int)prop.sharedMemPerBlockOptin;
               m->smem_per_sm = (int)prop.sharedMemPerMultiprocessor; }
    }
    if (rc) { pyjac_mech_destroy(m); return rc; }
    *out = m;
    return PYJAC_OK;
}

void pyjac_mech_destroy(pyjac_mech* m)
{
    if (!m) return;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_current == m) g_current = nullptr;
    }
    DeviceGuard guard(m->device);
    release_staging(m);
    for (int i = 0; i < 2; ++i) if (m->stream[i]) cudaStreamDestroy(m->stream[i]);
    for (void* p : m->dev_allocs) cudaFree(p);
    if (m->ws) cudaFree(m->ws);
    if (m->ws_ev) cudaEventDestroy(m->ws_ev);
    delete m;
}

int pyjac_mech_dims(const pyjac_mech* m, int dims[4])
{
    if (!m || !dims) return fail(PYJAC_EINVAL, "NULL argument");
    dims[0] = m->tb.nsp; dims[1] = m->tb.nr; dims[2] = m->tb.nrev; dims[3] = m->tb.npd;
    return PYJAC_OK;
}

int pyjac_mech_tune(pyjac_mech* m, int blocks_per_sm)
{
#ifndef PJ_DEV
    if (blocks_per_sm < 0) return fail(PYJAC_EINVAL, "bad argument");
#endif
    if (!m) return fail(PYJAC_EINVAL, "bad argument");
    m->user_bpsm = blocks_per_sm;
    m->bpsm[0] = m->bpsm[1] = m->bpsm[2] = m->bpsm[3] = 0;    // re-derive at next launch
    return PYJAC_OK;
}

int pyjac_mech_kernel_name(const pyjac_mech* m, int mode, char* buf, size_t len)
{
    if (!m || !buf || !len || mode < 0 || mode > 3) return fail(PYJAC_EINVAL, "bad argument");
    const void* fn = (mode == pj::M_JAC && m->has6) ? kernel6_for(m->plan6.gs, m->plan6.nt)
                                                     : kernel_for(m->plan.gs, mode, m->plan.nt, m->plan.wsg);
    if (!fn) return fail(PYJAC_EINVAL, "table blob holds no usable plan");
    const char* name = nullptr;
    CU(cudaFuncGetName(&name, fn));
    std::snprintf(buf, len, "%s", name ? name : "?");
    return PYJAC_OK;
}

long long pyjac_mech_launches(const pyjac_mech* m) { return m ? m->launches.load() : 0; }

int pyjac_eval_jacob_dev(pyjac_mech* m, int n, const double* d_pres, const double* d_y,
                         long long y_ss, long long y_sv, double* d_jac, int jac_layout,
                         long long jac_ld, void* stream)
{
    if (!m || n < 0 || (n && (!d_pres || !d_y || !d_jac))) return fail(PYJAC_EINVAL, "bad argument");
    if (jac_layout == PYJAC_JAC_STATE_FASTEST && (jac_ld < n || jac_ld >= (1LL << 28)))
        return fail(PYJAC_EINVAL, "jac_ld must be in [n, 2^28)");
    IO io{};
    io.n = n; io.pres = d_pres; io.y = d_y; io.y_ss = y_ss; io.y_sv = y_sv;
    io.jac = d_jac; io.jac_layout = jac_layout; io.jac_ld = jac_ld;
#ifdef PJ_DEV        // development builds only (tools/devbuild.sh NAME -DPJ_DEV): phase skipping, per-phase clocks, record checks
    if (const char* dbg = std::getenv("PYJAC_DEBUG_SKIP")) io.dbg_skip = std::atoi(dbg);
    if (const char* dbg = std::getenv("PYJAC_DEBUG_CLK")) io.dbg_clk = (long long*)std::strtoull(dbg, nullptr, 0);
#endif
    return launch(m, pj::M_JAC, io, (cudaStream_t)stream);
}


/* ---- factored Jacobian and its consumers (csrc/consumer.cuh) ---- */

int pyjac_factored_size(const pyjac_mech* m, int* nf, int* nnz)
{
    if (!m) return fail(PYJAC_EINVAL, "bad argument");
    if (nf) *nf = m->tb.nsp + 3 * (m->tb.nsp - 1) + m->fac.nnz;
    if (nnz) *nnz = m->fac.nnz;
    return PYJAC_OK;
}

int pyjac_factored_pattern(const pyjac_mech* m, int* rows, int* cols, double* ca, double* cb)
{
    if (!m) return fail(PYJAC_EINVAL, "bad argument");
    for (int p = 0; p < m->fac.nnz; ++p) {
        if (rows) rows[p] = m->fac_rows[p];
        if (cols) cols[p] = m->fac_cols[p];
    }
    for (int j = 0; j < m->tb.nsp; ++j) {
        if (ca) ca[j] = j ? m->colfac_h[2 * j] : 0.0;
        if (cb) cb[j] = j ? m->colfac_h[2 * j + 1] : 0.0;
    }
    return PYJAC_OK;
}

static int fac_layout_ok(const pyjac_mech* m, int n, int layout, long long ld)
{
    if (layout == PYJAC_JAC_STATE_FASTEST && (ld < n || ld >= (1LL << 28))) return fail(PYJAC_EINVAL, "ld must be in [n, 2^28)");
    if (layout != PYJAC_JAC_STATE_FASTEST && layout != PYJAC_JAC_STATE_MAJOR) return fail(PYJAC_EINVAL, "bad layout");
    (void)m;
    return PYJAC_OK;
}

int pyjac_eval_jacob_factored_dev(pyjac_mech* m, int n, const double* d_pres, const double* d_y,
                                  long long y_ss, long long y_sv, double* d_fac, int fac_layout,
                                  long long fac_ld, void* stream)
{
    if (!m || n < 0 || (n && (!d_pres || !d_y || !d_fac))) return fail(PYJAC_EINVAL, "bad argument");
    if (int rc = fac_layout_ok(m, n, fac_layout, fac_ld)) return rc;
    IO io{};
    io.n = n; io.pres = d_pres; io.y = d_y; io.y_ss = y_ss; io.y_sv = y_sv;
    io.jac = d_fac; io.jac_layout = fac_layout; io.jac_ld = fac_ld;
    return launch(m, pj::M_FACT, io, (cudaStream_t)stream);
}

int pyjac_jvp_dev(pyjac_mech* m, int n, const double* d_fac, int fac_layout, long long fac_ld,
                  const double* d_v, long long v_ss, long long v_sv,
                  double* d_out, long long o_ss, long long o_sv, void* stream)
{
    if (!m || n < 0 || (n && (!d_fac || !d_v || !d_out))) return fail(PYJAC_EINVAL, "bad argument");
    if (int rc = fac_layout_ok(m, n, fac_layout, fac_ld)) return rc;
    if (!n) return PYJAC_OK;
    DeviceGuard guard(m->device);
    pjc::Fac f = m->fac;
    f.fac = d_fac; f.sf = fac_layout == PYJAC_JAC_STATE_FASTEST; f.ld = fac_ld;
    pjc::k_jvp<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(f, n, d_v, v_ss, v_sv, d_out, o_ss, o_sv);
    CU(cudaGetLastError());
    ++m->launches;
    return PYJAC_OK;
}

int pyjac_newton_solve_dev(pyjac_mech* m, int n, const double* d_fac, int fac_layout, long long fac_ld,
                           double gamma, const double* d_gamma,
                           const double* d_rhs, long long r_ss, long long r_sv,
                           double* d_x, long long x_ss, long long x_sv, int* d_info, void* stream)
{
    if (!m || n < 0 || (n && (!d_fac || !d_rhs || !d_x))) return fail(PYJAC_EINVAL, "bad argument");
    if (int rc = fac_layout_ok(m, n, fac_layout, fac_ld)) return rc;
    if (!n) return PYJAC_OK;
    DeviceGuard guard(m->device);
    const int nsp = m->tb.nsp;
    const int ldm = nsp | 1;                                   // odd: row accesses spread over the banks
    const size_t per_warp = ((size_t)ldm * nsp + 3 * (size_t)nsp) * 8;
    if (per_warp > (size_t)m->smem_optin)
        return fail(PYJAC_ETOOBIG, "the Newton matrix of this mechanism does not fit in shared memory");
    const int wpb = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)m->smem_optin / per_warp));
    const size_t bytes = per_warp * wpb;
    const int rpl = (nsp + 31) / 32;                           // rows per lane
    const void* fn = rpl == 1 ? (const void*)pjc::k_newton<1> : rpl == 2 ? (const void*)pjc::k_newton<2>
                   : rpl == 3 ? (const void*)pjc::k_newton<3> : rpl == 4 ? (const void*)pjc::k_newton<4>
                   : rpl == 5 ? (const void*)pjc::k_newton<5> : rpl == 6 ? (const void*)pjc::k_newton<6> : nullptr;
    if (!fn) return fail(PYJAC_ETOOBIG, "the Newton matrix of this mechanism does not fit in shared memory");
    int rc = ensure_dyn_smem(fn, m->device, bytes, false);
    if (rc) return rc;
    const long long want = ((long long)n + wpb - 1) / wpb;
    int grid = (int)std::min<long long>(want, (long long)m->sm_count);
    pjc::Fac f = m->fac;
    f.fac = d_fac; f.sf = fac_layout == PYJAC_JAC_STATE_FASTEST; f.ld = fac_ld;
    int ldm_ = ldm;
    void* args[] = {(void*)&f, (void*)&n, (void*)&gamma, (void*)&d_gamma, (void*)&d_rhs, (void*)&r_ss, (void*)&r_sv,
                    (void*)&d_x, (void*)&x_ss, (void*)&x_sv, (void*)&d_info, (void*)&ldm_};
    CU(cudaLaunchKernel(fn, dim3(grid), dim3(wpb * 32), args, bytes, (cudaStream_t)stream));
    CU(cudaGetLastError());
    ++m->launches;
    return PYJAC_OK;
}

static int dydt_dev(pyjac_mech* m, int n, const double* d_var, const double* d_y, long long y_ss, long long y_sv,
                    double* d_dy, long long o_ss, long long o_sv, int conv, void* stream)
{
    if (!m || n < 0 || (n && (!d_var || !d_y || !d_dy))) return fail(PYJAC_EINVAL, "bad argument");
    IO io{};
    io.n = n; io.pres = d_var; io.y = d_y; io.y_ss = y_ss; io.y_sv = y_sv;
    io.dy = d_dy; io.dy_ss = o_ss; io.dy_sv = o_sv;
    io.conv = conv;
    return launch(m, pj::M_DYDT, io, (cudaStream_t)stream);
}

int pyjac_dydt_dev(pyjac_mech* m, int n, const double* d_pres, const double* d_y,
                   long long y_ss, long long y_sv, double* d_dy, long long o_ss,
                   long long o_sv, void* stream)
{
    return dydt_dev(m, n, d_pres, d_y, y_ss, y_sv, d_dy, o_ss, o_sv, 0, stream);
}

int pyjac_dydt_conv_dev(pyjac_mech* m, int n, const double* d_rho, const double* d_y,
                        long long y_ss, long long y_sv, double* d_dy, long long o_ss,
                        long long o_sv, void* stream)
{
    return dydt_dev(m, n, d_rho, d_y, y_ss, y_sv, d_dy, o_ss, o_sv, 1, stream);
}

int pyjac_mech_set_conv(pyjac_mech* m, int conv)
{
    if (!m) return fail(PYJAC_EINVAL, "bad argument");
    m->conv = conv ? 1 : 0;
    return PYJAC_OK;
}

// Finite-difference Jacobian of dydt on the device: the independent self-check of eval_jacob
// (the reference builds the same comparison from performance_tester/fd_jacob.cu).
int pyjac_fd_jacob_dev(pyjac_mech* m, int n, const double* d_pres, const double* d_y, double* d_jac,
                       int order, double r_cap, void* stream)
{
    if (!m || n < 0 || (n && (!d_pres || !d_y || !d_jac))) return fail(PYJAC_EINVAL, "bad argument");
    if (order != 1 && order != 2 && order != 4 && order != 6) return fail(PYJAC_EINVAL, "order must be 1, 2, 4 or 6");
    if (!n) return PYJAC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(m->device);
    const int nsp = m->tb.nsp;
    const size_t row = (size_t)n * nsp;
    double* w = nullptr;                     // dy0 | ytmp | dy | r
    CU(cudaMalloc(&w, (3 * row + (size_t)n) * sizeof(double)));
    double *dy0 = w, *ytmp = w + row, *dy = w + 2 * row, *r = w + 3 * row;
    static const double xs[4][6] = {{1}, {-1, 1}, {-2, -1, 1, 2}, {-3, -2, -1, 1, 2, 3}};
    static const double ws[4][6] = {{1}, {-0.5, 0.5}, {1.0 / 12, -2.0 / 3, 2.0 / 3, -1.0 / 12},
                                    {-1.0 / 60, 3.0 / 20, -3.0 / 4, 3.0 / 4, -3.0 / 20, 1.0 / 60}};
    const int oi = order == 1 ? 0 : order == 2 ? 1 : order == 4 ? 2 : 3;
    int rc = pyjac_dydt_dev(m, n, d_pres, d_y, 1, n, dy0, 1, n, st);
    cudaError_t ce = cudaMemcpyAsync(ytmp, d_y, row * sizeof(double), cudaMemcpyDeviceToDevice, st);
    const int tb_ = 256;
    const unsigned gs_ = (unsigned)((n + tb_ - 1) / tb_), ga_ = (unsigned)((row + tb_ - 1) / tb_);
    for (int j = 0; j < nsp && !rc && ce == cudaSuccess; ++j) {
        for (int k = 0; k < order && !rc; ++k) {
            pj::k_fd_step<<<gs_, tb_, 0, st>>>(n, nsp, j, xs[oi][k], r_cap, d_y, dy0, ytmp, r);
            rc = pyjac_dydt_dev(m, n, d_pres, ytmp, 1, n, dy, 1, n, st);
            pj::k_fd_accum<<<ga_, tb_, 0, st>>>(n, nsp, j, ws[oi][k], k == 0, dy, r, d_jac);
        }
        if (order == 1 && !rc)               // forward difference: (f(y + r) - f(y)) / r
            pj::k_fd_accum<<<ga_, tb_, 0, st>>>(n, nsp, j, -1.0, 0, dy0, r, d_jac);
        // restore row j of ytmp
        if (ce == cudaSuccess)
            ce = cudaMemcpyAsync(ytmp + (size_t)j * n, d_y + (size_t)j * n, (size_t)n * sizeof(double),
                                 cudaMemcpyDeviceToDevice, st);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(w);
    if (rc) return rc;
    if (ce != cudaSuccess) return fail(PYJAC_ECUDA, std::string("fd_jacob: ") + cudaGetErrorString(ce));
    return PYJAC_OK;
}

int pyjac_rates_dev(pyjac_mech* m, int n, const double* d_pres, const double* d_y,
                    long long y_ss, long long y_sv, double* d_conc, double* d_fwd,
                    double* d_rev, double* d_pres_mod, double* d_spec_rates, double* d_dy,
                    int o_state_fastest, long long o_ld, void* stream)
{
    if (!m || n < 0 || (n && (!d_pres || !d_y))) return fail(PYJAC_EINVAL, "bad argument");
    if (o_state_fastest && o_ld < n) return fail(PYJAC_EINVAL, "o_ld < n");
    IO io{};
    io.n = n; io.pres = d_pres; io.y = d_y; io.y_ss = y_ss; io.y_sv = y_sv;
    io.conc = d_conc; io.fwd = d_fwd; io.rev = d_rev; io.pm = d_pres_mod; io.sr = d_spec_rates;
    io.dy = d_dy; io.o_sf = o_state_fastest; io.o_ld = o_ld;
    return launch(m, pj::M_RATES, io, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- host-pointer batch API

// Streams row-major host states through two pinned/device staging slots:
// H2D(y, P) -> kernel -> D2H(result), chunk c+1 overlapping the D2H of chunk c.
static int host_stream(pyjac_mech* m, int n, const double* pres, const double* y, double* out, int kind)   // kind: 0 dydt, 1 Jacobian, 2 factored record
{
    if (!m || n < 0 || (n && (!pres || !y || !out))) return fail(PYJAC_EINVAL, "bad argument");
    if (!n) return PYJAC_OK;
    const int nsp = m->tb.nsp;
    const size_t in_w = (size_t)nsp + 1;                       // y row + pressure
    const size_t out_w = kind == 1 ? (size_t)nsp * nsp : kind == 2 ? (size_t)nsp + 3 * ((size_t)nsp - 1) + m->fac.nnz : (size_t)nsp;
    const size_t budget = (size_t)256 << 20;                   // bytes of output per chunk
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, budget / (out_w * 8)));
    int rc = ensure_staging(m, (size_t)chunk * in_w * 8, (size_t)chunk * in_w * 8, (size_t)chunk * out_w * 8);
    if (rc) return rc;
    int slot = 0;
    for (int s0 = 0; s0 < n; s0 += chunk, slot ^= 1) {
        const int cn = std::min(chunk, n - s0);
        cudaStream_t st = m->stream[slot];
        CU(cudaStreamSynchronize(st));                         // slot's previous D2H finished
        double* hp = m->h_pin[slot];
        std::memcpy(hp, y + (size_t)s0 * nsp, (size_t)cn * nsp * 8);
        std::memcpy(hp + (size_t)cn * nsp, pres + s0, (size_t)cn * 8);
        CU(cudaMemcpyAsync(m->d_in[slot], hp, (size_t)cn * in_w * 8, cudaMemcpyHostToDevice, st));
        const double* dy_ = m->d_in[slot];
        const double* dp_ = dy_ + (size_t)cn * nsp;
        if (kind == 1) rc = pyjac_eval_jacob_dev(m, cn, dp_, dy_, nsp, 1, m->d_out[slot], PYJAC_JAC_STATE_MAJOR, 0, st);
        else if (kind == 2) rc = pyjac_eval_jacob_factored_dev(m, cn, dp_, dy_, nsp, 1, m->d_out[slot], PYJAC_JAC_STATE_MAJOR, 0, st);
        else rc = pyjac_dydt_dev(m, cn, dp_, dy_, nsp, 1, m->d_out[slot], nsp, 1, st);
        if (rc) return rc;
        CU(cudaMemcpyAsync(out + (size_t)s0 * out_w, m->d_out[slot], (size_t)cn * out_w * 8,
                           cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < 2; ++i) CU(cudaStreamSynchronize(m->stream[i]));
    return PYJAC_OK;
}

int pyjac_eval_jacob_host(pyjac_mech* m, int n, const double* pres, const double* y, double* jac)
{
    return host_stream(m, n, pres, y, jac, 1);
}

int pyjac_eval_jacob_factored_host(pyjac_mech* m, int n, const double* pres, const double* y, double* fac)
{
    return host_stream(m, n, pres, y, fac, 2);
}

int pyjac_dydt_host(pyjac_mech* m, int n, const double* pres, const double* y, double* dy)
{
    return host_stream(m, n, pres, y, dy, 0);
}

int pyjac_set_mechanism(pyjac_mech* m)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_current = m;
    return PYJAC_OK;
}

int pyjac_register_tables(const void* blob, size_t len)
{
    if (!pjt::valid(blob, len)) return fail(PYJAC_EINVAL, "not a PJB200T1 table blob");
    std::lock_guard<std::mutex> lk(g_mu);
    g_reg_blob = blob;
    g_reg_len = len;
    return PYJAC_OK;
}

static pyjac_mech* current_or_null()
{
    pyjac_mech* m;
    const void* blob;
    size_t len;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        m = g_current;
        blob = g_reg_blob;
        len = g_reg_len;
    }
    if (m || !blob) return m;
    // first use of a library that carries its mechanism (pyjac_register_tables): load it on the
    // current device; concurrent first calls race benignly for g_current
    static std::mutex once;
    std::lock_guard<std::mutex> lk1(once);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_current) return g_current;
    }
    pyjac_mech* made = nullptr;
    if (pyjac_mech_create(blob, len, -1, &made) != PYJAC_OK) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    g_current = made;
    return made;
}

static pyjac_mech* current_or_die(const char* who)
{
    pyjac_mech* m = current_or_null();
    if (!m && g_reg_blob) {
        std::fprintf(stderr, "%s: %s\n", who, pyjac_last_error());
        std::exit(1);
    }
    if (!m) {
        std::fprintf(stderr, "%s: no mechanism selected (pyjac_set_mechanism)\n", who);
        std::exit(1);
    }
    return m;
}

static void die_on(int rc, const char* who)
{
    if (rc) {
        // the reference's entry points cannot report errors: cudaErrorCheck prints and exits
        // (mech_auxiliary.py:424-436)
        std::fprintf(stderr, "%s: %s\n", who, pyjac_last_error());
        std::exit(1);
    }
}

int pyjac_cu_init(int num)
{
    pyjac_mech* m = current_or_null();
    if (!m) return fail(PYJAC_EINVAL, "no mechanism selected (pyjac_set_mechanism)");
    if (num <= 0) return fail(PYJAC_EINVAL, "num must be positive");
    g_cu_num = num;
    return (num + 7) / 8 * 8;
}

// device chunk of state-fastest arrays with leading dimension ld
static int cu_run_impl(pyjac_mech* m, int num, const double* pres, const double* mass_frac,
                       double* conc, double* fwd, double* rev, double* pmod, double* sr,
                       double* dy, double* jac)
{
    const Tables& t = m->tb;
    const int nsp = t.nsp;
    const size_t widths[7] = {(size_t)nsp, (size_t)t.nr, (size_t)t.nrev, (size_t)t.npd, (size_t)nsp,
                              (size_t)nsp, (size_t)nsp * nsp};
    double* outs[7] = {conc, fwd, rev, pmod, sr, dy, jac};
    size_t wsum = 0;
    for (int i = 0; i < 7; ++i) if (outs[i]) wsum += widths[i];
    const size_t budget = (size_t)512 << 20;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)num, budget / ((wsum + nsp + 1) * 8)));
    const int ld = (chunk + 7) / 8 * 8;
    int rc = ensure_staging(m, 16, (size_t)ld * (nsp + 1) * 8, (size_t)ld * std::max<size_t>(wsum, 1) * 8);
    if (rc) return rc;
    cudaStream_t st = m->stream[0];
    for (int s0 = 0; s0 < num; s0 += chunk) {
        const int cn = std::min(chunk, num - s0);
        double* d_y = m->d_in[0];
        double* d_p = d_y + (size_t)ld * nsp;
        CU(cudaMemcpy2DAsync(d_y, (size_t)ld * 8, mass_frac + s0, (size_t)num * 8, (size_t)cn * 8, nsp,
                             cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_p, pres + s0, (size_t)cn * 8, cudaMemcpyHostToDevice, st));
        double* d_o[7];
        size_t off = 0;
        for (int i = 0; i < 7; ++i) {
            d_o[i] = outs[i] ? m->d_out[0] + off : nullptr;
            if (outs[i]) off += widths[i] * ld;
        }
        if (d_o[0] || d_o[1] || d_o[2] || d_o[3] || d_o[4] || d_o[5]) {
            rc = pyjac_rates_dev(m, cn, d_p, d_y, 1, ld, d_o[0], d_o[1], d_o[2], d_o[3], d_o[4], d_o[5], 1, ld, st);
            if (rc) return rc;
        }
        if (d_o[6]) {
            rc = pyjac_eval_jacob_dev(m, cn, d_p, d_y, 1, ld, d_o[6], PYJAC_JAC_STATE_FASTEST, ld, st);
            if (rc) return rc;
        }
        for (int i = 0; i < 7; ++i)
            if (outs[i] && widths[i])
                CU(cudaMemcpy2DAsync(outs[i] + s0, (size_t)num * 8, d_o[i], (size_t)ld * 8, (size_t)cn * 8,
                                     widths[i], cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return PYJAC_OK;
}

void pyjac_cu_run(int num, int padded, const double* pres, const double* mass_frac,
                  double* conc, double* fwd_rxn_rates, double* rev_rxn_rates,
                  double* pres_mod, double* spec_rates, double* dy, double* jac)
{
    (void)padded;
    pyjac_mech* m = current_or_die("run");
    die_on(cu_run_impl(m, num, pres, mass_frac, conc, fwd_rxn_rates, rev_rxn_rates, pres_mod,
                       spec_rates, dy, jac), "run");
}

void pyjac_cu_cleanup(void)
{
    pyjac_mech* m;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        m = g_current;
    }
    if (m) { DeviceGuard guard(m->device); release_staging(m); }
    g_cu_num = 0;
    std::vector<Slot*> slots;
    {
        std::lock_guard<std::mutex> lk(g_slot_mu);
        slots.swap(g_slots);
    }
    for (Slot* s : slots) { DeviceGuard guard(s->device); slot_free(s); }
}

// ---------------------------------------------------------------- scalar API (batch of 1)
// Every call borrows its own staging slot (stream, pinned and device buffers): re-entrant and safe to
// call from concurrent host threads, as the reference's stack-only functions are (tester.c.in:24-29).

static int scalar_state(pyjac_mech* m, double pres, const double* y, double* out, bool jac)
{
    const int nsp = m->tb.nsp;
    const size_t in_b = ((size_t)nsp + 1) * 8, out_b = (jac ? (size_t)nsp * nsp : (size_t)nsp) * 8;
    DeviceGuard guard(m->device);
    SlotLease L;
    int rc = slot_acquire(m->device, std::max(in_b, out_b), in_b, out_b, &L.s);
    if (rc) return rc;
    Slot* s = L.s;
    std::memcpy(s->h, y, (size_t)nsp * 8);
    s->h[nsp] = pres;
    CU(cudaMemcpyAsync(s->d_in, s->h, in_b, cudaMemcpyHostToDevice, s->st));
    if (jac && m->conv) return fail(PYJAC_EINVAL, "eval_jacob has no constant-volume form (the reference emits none either)");
    if (jac) rc = pyjac_eval_jacob_dev(m, 1, s->d_in + nsp, s->d_in, nsp, 1, s->d_out, PYJAC_JAC_STATE_MAJOR, 0, s->st);
    else rc = dydt_dev(m, 1, s->d_in + nsp, s->d_in, nsp, 1, s->d_out, nsp, 1, m->conv, s->st);
    if (rc) { cudaStreamSynchronize(s->st); return rc; }
    CU(cudaMemcpyAsync(s->h, s->d_out, out_b, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    std::memcpy(out, s->h, out_b);
    return PYJAC_OK;
}

void eval_jacob(const double t, const double pres, const double* y, double* jac)
{
    (void)t;
    pyjac_mech* m = current_or_die("eval_jacob");
    die_on(scalar_state(m, pres, y, jac, true), "eval_jacob");
}

void dydt(const double t, const double pres, const double* y, double* dy)
{
    (void)t;
    pyjac_mech* m = current_or_die("dydt");
    die_on(scalar_state(m, pres, y, dy, false), "dydt");
}

// one state through the M_RATES kernel; in_conc selects [T, C...] input
static int scalar_rates(pyjac_mech* m, double T, double pres, const double* in, int n_in, bool in_conc,
                        double* conc, double* fwd, double* rev, double* pmod, double* extra3)
{
    const Tables& t = m->tb;
    const int nsp = t.nsp;
    const size_t in_d = (size_t)nsp + 2, n_out = (size_t)nsp + t.nr + t.nrev + t.npd, out_d = n_out + 4;
    DeviceGuard guard(m->device);
    SlotLease L;
    int rc = slot_acquire(m->device, std::max(in_d, out_d) * 8, in_d * 8, out_d * 8, &L.s);
    if (rc) return rc;
    Slot* s = L.s;
    double* hp = s->h;
    hp[0] = T;
    std::memcpy(hp + 1, in, (size_t)n_in * 8);
    hp[nsp + 1] = pres;
    CU(cudaMemcpyAsync(s->d_in, hp, in_d * 8, cudaMemcpyHostToDevice, s->st));
    double* d = s->d_out;
    IO io{};
    io.n = 1; io.pres = s->d_in + nsp + 1; io.y = s->d_in; io.y_ss = 0; io.y_sv = 1;
    io.in_conc = in_conc ? 1 : 0;
    io.conc = d; io.fwd = d + nsp; io.rev = d + nsp + t.nr; io.pm = d + nsp + t.nr + t.nrev;
    io.scal3 = d + n_out;
    rc = launch(m, pj::M_RATES, io, s->st);
    if (rc) { cudaStreamSynchronize(s->st); return rc; }
    CU(cudaMemcpyAsync(hp, d, out_d * 8, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    if (conc) std::memcpy(conc, hp, (size_t)nsp * 8);
    if (fwd) std::memcpy(fwd, hp + nsp, (size_t)t.nr * 8);
    if (rev) std::memcpy(rev, hp + nsp + t.nr, (size_t)t.nrev * 8);
    if (pmod) std::memcpy(pmod, hp + nsp + t.nr + t.nrev, (size_t)t.npd * 8);
    if (extra3) std::memcpy(extra3, hp + n_out, 3 * 8);
    return PYJAC_OK;
}

void eval_conc(const double T, const double pres, const double* mass_frac, double* y_N,
               double* mw_avg, double* rho, double* conc)
{
    pyjac_mech* m = current_or_die("eval_conc");
    const int nsp = m->tb.nsp;
    double s3[3];
    die_on(scalar_rates(m, T, pres, mass_frac, nsp - 1, false, conc, nullptr, nullptr, nullptr, s3),
           "eval_conc");
    if (y_N) *y_N = s3[0];
    if (mw_avg) *mw_avg = s3[1];
    if (rho) *rho = s3[2];
}

void eval_rxn_rates(const double T, const double pres, const double* C, double* fwd_rxn_rates,
                    double* rev_rxn_rates)
{
    pyjac_mech* m = current_or_die("eval_rxn_rates");
    die_on(scalar_rates(m, T, pres, C, m->tb.nsp, true, nullptr, fwd_rxn_rates, rev_rxn_rates, nullptr, nullptr),
           "eval_rxn_rates");
}

void get_rxn_pres_mod(const double T, const double pres, const double* C, double* pres_mod)
{
    pyjac_mech* m = current_or_die("get_rxn_pres_mod");
    die_on(scalar_rates(m, T, pres, C, m->tb.nsp, true, nullptr, nullptr, nullptr, pres_mod, nullptr),
           "get_rxn_pres_mod");
}

void eval_spec_rates(const double* fwd_rates, const double* rev_rates, const double* pres_mod,
                     double* sp_rates, double* dy_N)
{
    pyjac_mech* m = current_or_die("eval_spec_rates");
    const Tables& t = m->tb;
    const size_t in_d = (size_t)t.nr + t.nrev + t.npd + 1, out_d = (size_t)t.nsp;
    auto run = [&]() -> int {
        DeviceGuard guard(m->device);
        SlotLease L;
        int rc = slot_acquire(m->device, std::max(in_d, out_d) * 8, in_d * 8, out_d * 8, &L.s);
        if (rc) return rc;
        Slot* s = L.s;
        double* hp = s->h;
        std::memcpy(hp, fwd_rates, (size_t)t.nr * 8);
        if (t.nrev) std::memcpy(hp + t.nr, rev_rates, (size_t)t.nrev * 8);
        if (t.npd) std::memcpy(hp + t.nr + t.nrev, pres_mod, (size_t)t.npd * 8);
        CU(cudaMemcpyAsync(s->d_in, hp, in_d * 8, cudaMemcpyHostToDevice, s->st));
        pj::k_spec_rates<<<(t.nsp + 63) / 64, 64, 0, s->st>>>(t, s->d_in, s->d_in + t.nr, s->d_in + t.nr + t.nrev, s->d_out);
        CU(cudaGetLastError());
        ++m->launches;
        CU(cudaMemcpyAsync(hp, s->d_out, out_d * 8, cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
        std::memcpy(sp_rates, hp, (size_t)(t.nsp - 1) * 8);
        *dy_N = hp[t.nsp - 1];
        return PYJAC_OK;
    };
    die_on(run(), "eval_spec_rates");
}

// eval_h / eval_u / eval_cv / eval_cp (chem_utils.h of the emitted library, rate_subs.py:1581-1608): NSP
// mass-based values for one temperature
static void scalar_thermo(const char* who, int what, double T, double* out)
{
    pyjac_mech* m = current_or_die(who);
    const Tables& t = m->tb;
    auto run = [&]() -> int {
        DeviceGuard guard(m->device);
        SlotLease L;
        int rc = slot_acquire(m->device, (size_t)t.nsp * 8, 8, (size_t)t.nsp * 8, &L.s);
        if (rc) return rc;
        Slot* s = L.s;
        pj::k_thermo<<<(t.nsp + 63) / 64, 64, 0, s->st>>>(t, T, what, s->d_out);
        CU(cudaGetLastError());
        ++m->launches;
        CU(cudaMemcpyAsync(s->h, s->d_out, (size_t)t.nsp * 8, cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
        std::memcpy(out, s->h, (size_t)t.nsp * 8);
        return PYJAC_OK;
    };
    die_on(run(), who);
}

void eval_h(const double T, double* h) { scalar_thermo("eval_h", pj::TH_H, T, h); }
void eval_u(const double T, double* u) { scalar_thermo("eval_u", pj::TH_U, T, u); }
void eval_cv(const double T, double* cv) { scalar_thermo("eval_cv", pj::TH_CV, T, cv); }
void eval_cp(const double T, double* cp) { scalar_thermo("eval_cp", pj::TH_CP, T, cp); }

// apply_mask / apply_reverse_mask (mech_auxiliary.py:188-206): the species permutation that moves the last
// species to the end, applied to / undone on an array of NSP mass fractions in place.  Host-side data
// movement of NSP doubles, as in the reference (read_initial_conditions.c:29).
void apply_mask(double* y_specs)
{
    pyjac_mech* m = current_or_die("apply_mask");
    const int nsp = m->tb.nsp;
    std::vector<double> tmp(y_specs, y_specs + nsp);
    for (int i = 0; i < nsp; ++i) y_specs[i] = tmp[m->fwd_map[i]];
}

void apply_reverse_mask(double* y_specs)
{
    pyjac_mech* m = current_or_die("apply_reverse_mask");
    const int nsp = m->tb.nsp;
    std::vector<double> tmp(y_specs, y_specs + nsp);
    for (int i = 0; i < nsp; ++i) y_specs[i] = tmp[m->back_map[i]];
}

}  // extern "C"

// ---------------------------------------------------------------- the reference's batched GPU host API
// pyjac/pywrap/pyjacob.cuh:6-10 declares init / run / cleanup with C++ linkage (the header is included
// by pyjacob.cu and by the Cython C++ extension, pyjacob_cuda_wrapper.pyx:5-10): the same three symbols,
// same signatures, so that a build of the reference's cu_pyjacob wrapper links against this library.
int init(int num)
{
    const int padded = pyjac_cu_init(num);
    if (padded < 0) {
        std::fprintf(stderr, "init: %s\n", pyjac_last_error());
        std::exit(1);
    }
    return padded;
}

void run(int num, int padded, const double* pres, const double* mass_frac, double* conc, double* fwd_rxn_rates,
         double* rev_rxn_rates, double* pres_mod, double* spec_rates, double* dy, double* jac)
{
    pyjac_cu_run(num, padded, pres, mass_frac, conc, fwd_rxn_rates, rev_rxn_rates, pres_mod, spec_rates, dy, jac);
}

void cleanup() { pyjac_cu_cleanup(); }
