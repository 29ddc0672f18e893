// qt_reference_api.cpp — the reference's harness-level operators (main.cuh:61-70) over the C ABI
// (include/qtesla_b200_reference_api.h).  Host-only C++: no CUDA headers, only qt_* calls; built into
// libqtesla_b200.so so that a maintainer links one library.
//
// State: a qt_ref_session (parameter set, batch, device, context) is an explicit handle.  The six functions with
// the reference's own signatures cannot carry one — main.cuh:61-70 fixes their argument lists — so they run on ONE
// default session, which qt_ref_configure replaces and a mutex serialises; everything else takes the handle.
#include "../../include/qtesla_b200_reference_api.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <new>

#include "../../include/qtesla_b200.h"

struct qt_ref_session {
    int set = QT_SET_III, device = 0, keep = 0;
    uint64_t batch = 2;  // BATCH of main.cuh:7
    qt_ctx* ctx = nullptr;
    qt_params p{};
};

namespace {

const char* const kTitle[] = {"test_NTT_negacyclic Stockham GPU", "test_NTT_negacyclic GS-CT GPU", "test_NTT_negacyclic CT-CT GPU",
                              "test_NTT_negacyclic GS-GS GPU", "test_NTT_negacyclic CT-GS GPU", "test_nussbaumer GPU"};
const char* const kLabel[] = {"Stockham GPU", "GS-CT GPU", "CT-CT GPU", "GS-GS GPU", "CT-GS GPU", "Nussbaumer GPU"};

int session_run(qt_ref_session* s, int driver, uint32_t* x, uint32_t* y, uint32_t* out) {
    if (!s || driver < QT_REF_STOCKHAM || driver > QT_REF_NUSSBAUMER || !x || !y || !out) return QT_ERR_BAD_ARG;
    if (!s->ctx) {
        const int rc = qt_create(s->set, s->device, &s->ctx);
        if (rc) return rc;
    }
    const uint64_t words = s->batch * s->p.n;
    if (!s->keep)
        for (uint64_t i = 0; i < words; i++) { x[i] = 1; y[i] = 1; }  // NTT.cu:2010, 2099, 2183, 2273, 2360
    printf("\n========================\n");
    printf("%s. Batch Size is %llu", kTitle[driver], (unsigned long long)s->batch);
    printf("\n========================\n");
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = driver == QT_REF_NUSSBAUMER ? qt_nussbaumer_host(s->ctx, x, y, out, s->batch, QT_RING_2P32M1)
                                               : qt_polymul_host(s->ctx, x, y, out, s->batch);  // H2D + kernel + D2H, like NTT.cu:2123-2164
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc) return rc;
    printf("Performance GPU %s \n Time\t\t: % .4f ms. \nThroughput\t: %.2f Multiplications per second\n", kLabel[driver], ms,
           (double)s->batch / ms * 1000.0);
    if (!s->keep) {  // the reference's DEBUG dump (NTT.cu:2084-2090)
        printf("z: ");
        for (uint64_t i = 0; i < words; i++) {
            if (i % s->p.n == 0) printf("\n\n");
            printf("%u ", out[i]);
        }
    }
    return 0;
}

std::mutex g_default_mutex;
qt_ref_session* g_default = nullptr;  // the session behind the six handle-less reference signatures

void run_default(int driver, uint32_t* x, uint32_t* y, uint32_t* out) {
    std::lock_guard<std::mutex> lk(g_default_mutex);
    int rc = 0;
    if (!g_default) rc = qt_ref_open(QT_SET_III, 2, 0, &g_default);  // the reference's compile-time configuration
    if (!rc) rc = session_run(g_default, driver, x, y, out);
    if (rc) {  // the reference's drivers return void: nothing to hand an error to
        fprintf(stderr, "qtesla_b200: %s\n", qt_error_string(rc));
        exit(1);
    }
}

}  // namespace

extern "C" {

int qt_ref_open(int set, uint64_t batch, int device, qt_ref_session** out) {
    if (!out) return QT_ERR_BAD_ARG;
    *out = nullptr;
    qt_params p;
    const int rc = qt_get_params(set, &p);
    if (rc) return rc;
    if (batch == 0) return QT_ERR_BAD_ARG;
    qt_ref_session* s = new (std::nothrow) qt_ref_session();
    if (!s) return QT_ERR_NOMEM;
    s->set = set; s->batch = batch; s->device = device; s->p = p;
    *out = s;  // the context is created by the first run (so that opening needs no GPU)
    return 0;
}

int qt_ref_close(qt_ref_session* s) {
    if (!s) return 0;
    if (s->ctx) qt_destroy(s->ctx);
    delete s;
    return 0;
}

int qt_ref_session_keep_operands(qt_ref_session* s, int keep) {
    if (!s) return QT_ERR_BAD_ARG;
    s->keep = keep;
    return 0;
}

int qt_ref_run(qt_ref_session* s, int driver, uint32_t* x, uint32_t* y, uint32_t* out) { return session_run(s, driver, x, y, out); }

int qt_ref_configure(int set, uint64_t batch, int device) {
    qt_ref_session* fresh = nullptr;
    const int rc = qt_ref_open(set, batch, device, &fresh);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_default_mutex);
    if (g_default) {
        fresh->keep = g_default->keep;
        qt_ref_close(g_default);
    }
    g_default = fresh;
    return 0;
}

void qt_ref_keep_operands(int keep) {
    std::lock_guard<std::mutex> lk(g_default_mutex);
    if (!g_default && qt_ref_open(QT_SET_III, 2, 0, &g_default)) return;
    g_default->keep = keep;
}

void test_NTT_Stockham_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t* Z, uint32_t*,
                                uint32_t*, uint32_t, uint32_t, uint32_t) {
    run_default(QT_REF_STOCKHAM, x, y, Z ? Z : z);  // the Stockham driver leaves its result in Z (NTT.cu:2078)
}
void test_NTT_GS_CT_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*,
                             uint32_t, uint32_t, uint32_t) {
    run_default(QT_REF_GS_CT, x, y, z);
}
void test_NTT_CT_CT_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*,
                             uint32_t, uint32_t, uint32_t) {
    run_default(QT_REF_CT_CT, x, y, z);
}
void test_NTT_GS_GS_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*,
                             uint32_t, uint32_t, uint32_t) {
    run_default(QT_REF_GS_GS, x, y, z);
}
void test_NTT_CT_GS_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*,
                             uint32_t, uint32_t, uint32_t) {
    run_default(QT_REF_CT_GS, x, y, z);
}
void test_nussbaumer(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t*) {
    run_default(QT_REF_NUSSBAUMER, x, y, z);
}

}  // extern "C"
