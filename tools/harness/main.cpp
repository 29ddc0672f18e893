// main.cpp — Linux re-creation of the reference's command-line harness (main.cu:12-230) on top of
// libqtesla_b200.so.  Same flags: -speedgpu k (k = 2..6, 8 as in main.cu:197-225; 9 = Nussbaumer on the
// GPU), -r seed (parsed, unused — as in the reference, main.cu:89-92).  Added: -set I|III|p-I|p-III,
// -batch B (the reference's compile-time BATCH/NTTSIZE/P), -device d, -quiet (no result dump).
// The reference's -cpu / -speedcpu options run CPU code; this product has no CPU path — they are
// answered with a pointer to the test oracle and exit code 2.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/qtesla_b200.h"
#include "../../include/qtesla_b200_reference_api.h"

static void help_message() {
    printf("-speedgpu \t 2: Stockham\t 3: GS-CT\t 4: CT-CT\t 5: GS-GS\t 6: CT-GS\t 8: GS-CT x5 + CT-GS x5\t 9: Nussbaumer\n");
    printf("-set I|III|p-I|p-III   -batch B   -device d   -quiet   -r seed\n");
    printf("-cpu / -speedcpu: CPU paths are not part of this engine (see oracle/ for the CPU oracle)\n");
}

int main(int argc, char** argv) {
    int set = QT_SET_III, device = 0, gpu_option = -1, quiet = 0;
    unsigned long long batch = 2;
    unsigned seed = 0;
    if (argc < 3) { help_message(); return -1; }
    for (int i = 1; i < argc;) {
        if (!strcmp(argv[i], "-speedgpu") && i + 1 < argc) { gpu_option = atoi(argv[i + 1]); i += 2; }
        else if (!strcmp(argv[i], "-r") && i + 1 < argc) { seed = (unsigned)atoi(argv[i + 1]); i += 2; }
        else if (!strcmp(argv[i], "-batch") && i + 1 < argc) { batch = strtoull(argv[i + 1], nullptr, 10); i += 2; }
        else if (!strcmp(argv[i], "-device") && i + 1 < argc) { device = atoi(argv[i + 1]); i += 2; }
        else if (!strcmp(argv[i], "-quiet")) { quiet = 1; i += 1; }
        else if (!strcmp(argv[i], "-set") && i + 1 < argc) {
            const char* s = argv[i + 1];
            set = !strcmp(s, "I") ? QT_SET_I : !strcmp(s, "III") ? QT_SET_III : !strcmp(s, "p-I") ? QT_SET_P_I
                : !strcmp(s, "p-III") ? QT_SET_P_III : -1;
            i += 2;
        } else if (!strcmp(argv[i], "-cpu") || !strcmp(argv[i], "-speedcpu")) {
            fprintf(stderr, "%s: the CPU paths of the reference are not part of this engine; the CPU oracle lives in oracle/\n", argv[i]);
            return 2;
        } else { help_message(); return -1; }
    }
    (void)seed;
    qt_params p;
    int rc = qt_get_params(set, &p);
    if (rc || batch == 0) { fprintf(stderr, "bad -set / -batch\n"); return -1; }
    printf("NTT Parameters==> NTTSIZE: %u P: %u omega: %u psi: %u Ni: %u BATCH: %llu\n", p.n, p.q, p.omega, p.psi, p.n_inv, batch);
    rc = qt_ref_configure(set, batch, device);
    if (rc) { fprintf(stderr, "%s\n", qt_error_string(rc)); return 1; }
    qt_ref_keep_operands(0);
    const size_t words = (size_t)batch * p.n;
    uint32_t *x, *y, *z;
    // pinned host buffers (the reference mallocs pageable ones, main.cu:106-108)
    if (qt_host_alloc(words * 4, (void**)&x) || qt_host_alloc(words * 4, (void**)&y) || qt_host_alloc(words * 4, (void**)&z)) {
        fprintf(stderr, "host allocation failed (no CUDA device?)\n");
        return 1;
    }
    if (quiet) { for (size_t i = 0; i < words; i++) { x[i] = 1; y[i] = 1; } qt_ref_keep_operands(1); }
    printf("\n\n========================\nSpeed Test\n========================\n");
    switch (gpu_option) {
    case 2: test_NTT_Stockham_nega_gpu(x, y, z, 0, 0, z, 0, 0, 0, 0, p.n_inv); break;
    case 3: test_NTT_GS_CT_nega_gpu(x, y, z, 0, 0, 0, 0, 0, 0, 0, p.n_inv); break;
    case 4: test_NTT_CT_CT_nega_gpu(x, y, z, 0, 0, 0, 0, 0, 0, 0, p.n_inv); break;
    case 5: test_NTT_GS_GS_nega_gpu(x, y, z, 0, 0, 0, 0, 0, 0, 0, p.n_inv); break;
    case 6: test_NTT_CT_GS_nega_gpu(x, y, z, 0, 0, 0, 0, 0, 0, 0, p.n_inv); break;
    case 8:
        for (int i = 0; i < 5; i++) test_NTT_GS_CT_nega_gpu(x, y, z, 0, 0, 0, 0, 0, 0, 0, p.n_inv);
        for (int i = 0; i < 5; i++) test_NTT_CT_GS_nega_gpu(x, y, z, 0, 0, 0, 0, 0, 0, 0, p.n_inv);
        break;
    case 9: test_nussbaumer(x, y, z, 0, 0, 0); break;
    default: help_message(); return -1;
    }
    printf("\n");
    qt_host_free(x); qt_host_free(y); qt_host_free(z);
    return 0;
}
