// qt_reference_api.cpp — the reference's harness-level operators (main.cuh:61-70) over the C ABI
// (include/qtesla_b200_reference_api.h).  Host-only C++: no CUDA headers, only qt_* calls; built into
// libqtesla_b200.so so that a maintainer links one library.
#include "../../include/qtesla_b200_reference_api.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../../include/qtesla_b200.h"

namespace {
int g_set = QT_SET_III, g_device = 0, g_keep = 0;
uint64_t g_batch = 2;  // BATCH of main.cuh:7
qt_ctx* g_ctx = nullptr;
qt_params g_p;

bool ensure() {
    if (g_ctx) return true;
    int rc = qt_get_params(g_set, &g_p);
    if (!rc) rc = qt_create(g_set, g_device, &g_ctx);
    if (rc) {
        fprintf(stderr, "qtesla_b200: %s\n", qt_error_string(rc));
        return false;
    }
    return true;
}

void run(const char* title, const char* label, uint32_t* x, uint32_t* y, uint32_t* out, int nuss) {
    if (!ensure()) exit(1);
    const uint64_t words = g_batch * g_p.n;
    if (!g_keep)
        for (uint64_t i = 0; i < words; i++) { x[i] = 1; y[i] = 1; }
    printf("\n========================\n");
    printf("%s. Batch Size is %llu", title, (unsigned long long)g_batch);
    printf("\n========================\n");
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = nuss ? qt_nussbaumer_host(g_ctx, x, y, out, g_batch, QT_RING_2P32M1)
                        : qt_polymul_host(g_ctx, x, y, out, g_batch);  // H2D + kernel + D2H, like NTT.cu:2123-2164
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc) {
        fprintf(stderr, "qtesla_b200: %s\n", qt_error_string(rc));
        exit(1);
    }
    printf("Performance GPU %s \n Time\t\t: % .4f ms. \nThroughput\t: %.2f Multiplications per second\n", label, ms,
           (double)g_batch / ms * 1000.0);
    if (!g_keep) {  // the reference's DEBUG dump (NTT.cu:2084-2090)
        printf("z: ");
        for (uint64_t i = 0; i < words; i++) {
            if (i % g_p.n == 0) printf("\n\n");
            printf("%u ", out[i]);
        }
    }
}
}  // namespace

extern "C" {

int qt_ref_configure(int set, uint64_t batch, int device) {
    qt_params p;
    int rc = qt_get_params(set, &p);
    if (rc) return rc;
    if (batch == 0) return QT_ERR_BAD_ARG;
    if (g_ctx) { qt_destroy(g_ctx); g_ctx = nullptr; }
    g_set = set; g_batch = batch; g_device = device;
    return 0;
}
void qt_ref_keep_operands(int keep) { g_keep = keep; }

void test_NTT_Stockham_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t* Z, uint32_t*,
                                uint32_t*, uint32_t, uint32_t, uint32_t) {
    run("test_NTT_negacyclic Stockham GPU", "Stockham GPU", x, y, Z ? Z : z, 0);
}
void test_NTT_GS_CT_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*,
                             uint32_t, uint32_t, uint32_t) {
    run("test_NTT_negacyclic GS-CT GPU", "GS-CT GPU", x, y, z, 0);
}
void test_NTT_CT_CT_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*,
                             uint32_t, uint32_t, uint32_t) {
    run("test_NTT_negacyclic CT-CT GPU", "CT-CT GPU", x, y, z, 0);
}
void test_NTT_GS_GS_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*,
                             uint32_t, uint32_t, uint32_t) {
    run("test_NTT_negacyclic GS-GS GPU", "GS-GS GPU", x, y, z, 0);
}
void test_NTT_CT_GS_nega_gpu(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t*,
                             uint32_t, uint32_t, uint32_t) {
    run("test_NTT_negacyclic CT-GS GPU", "CT-GS GPU", x, y, z, 0);
}
void test_nussbaumer(uint32_t* x, uint32_t* y, uint32_t* z, uint32_t*, uint32_t*, uint32_t*) {
    run("test_nussbaumer GPU", "Nussbaumer GPU", x, y, z, 1);
}

}  // extern "C"
