// pcie_ceiling.cu — the bare host<->device copy ceiling of the box for 1, 2, 4, 8 GPUs at once, in the
// traffic pattern of one end-to-end step of the hot path (per GPU: 512 MiB host->device for x and y and
// 256 MiB device->host for z, both directions in flight together).  One host thread per GPU; no kernel runs.
// This is the denominator of the `e2e` figure: qt_polymul_host cannot be faster than these copies.
//
//   nvcc -O3 -std=c++17 -o tools/pcie_ceiling tools/pcie_ceiling.cu && tools/pcie_ceiling [max_gpus]
//
// Prints one JSON object per line: the topology the driver / sysfs report, host memcpy bandwidth (what a
// staged pageable pipeline is bounded by), then one line per (gpus, allocation kind, direction, chunking).
// Allocation kinds: "default" cudaHostAlloc, "wc" write-combined (H2D source only), "numa" = the thread is
// bound to the CPUs of the GPU's NUMA node (sysfs) before it allocates and touches its buffers.
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

static const size_t MB = 1u << 20;
static const size_t IN_BYTES = 512 * MB, OUT_BYTES = 256 * MB;

struct Barrier {  // sense-reversing spin barrier
    std::atomic<int> count{0}, sense{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        const int s = sense.load();
        if (count.fetch_add(1) + 1 == n) { count = 0; sense = s ^ 1; }
        else while (sense.load() == s) std::this_thread::yield();
    }
};

static std::string read_file(const std::string& p) {
    FILE* f = fopen(p.c_str(), "r");
    if (!f) return "";
    char buf[4096];
    size_t k = fread(buf, 1, sizeof(buf) - 1, f);
    fclose(f);
    buf[k] = 0;
    while (k && (buf[k - 1] == '\n' || buf[k - 1] == ' ')) buf[--k] = 0;
    return buf;
}

static int gpu_numa_node(int dev, std::string* pci_out) {
    char pci[32];
    CK(cudaDeviceGetPCIBusId(pci, sizeof(pci), dev));
    for (char* c = pci; *c; c++) *c = (char)tolower(*c);
    *pci_out = pci;
    const std::string s = read_file(std::string("/sys/bus/pci/devices/") + pci + "/numa_node");
    return s.empty() ? -1 : atoi(s.c_str());
}

// "0-15,32-47" -> cpu_set_t
static bool parse_cpulist(const std::string& s, cpu_set_t* set) {
    CPU_ZERO(set);
    bool any = false;
    const char* p = s.c_str();
    while (*p) {
        char* e;
        long a = strtol(p, &e, 10), b = a;
        if (e == p) break;
        if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
        for (long c = a; c <= b; c++) { CPU_SET((int)c, set); any = true; }
        p = (*e == ',') ? e + 1 : e;
        if (*e != ',') break;
    }
    return any;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void run_case(int ngpus, const char* alloc, bool h2d, bool d2h, size_t chunk, int reps) {
    Barrier bar(ngpus);
    std::vector<double> per(ngpus, 0.0);
    double t_total = 0;
    std::vector<std::thread> th;
    for (int g = 0; g < ngpus; g++) {
        th.emplace_back([&, g]() {
            CK(cudaSetDevice(g));
            if (!strcmp(alloc, "numa")) {
                std::string pci;
                const int node = gpu_numa_node(g, &pci);
                cpu_set_t set;
                if (node >= 0 && parse_cpulist(read_file("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist"), &set))
                    sched_setaffinity(0, sizeof(set), &set);
            }
            char *h_in, *h_out, *d_in, *d_out;
            CK(cudaHostAlloc(&h_in, IN_BYTES, !strcmp(alloc, "wc") ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
            CK(cudaHostAlloc(&h_out, OUT_BYTES, cudaHostAllocDefault));
            memset(h_in, 1, IN_BYTES);
            memset(h_out, 1, OUT_BYTES);
            CK(cudaMalloc(&d_in, IN_BYTES));
            CK(cudaMalloc(&d_out, OUT_BYTES));
            cudaStream_t s1, s2;
            CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
            auto once = [&]() {
                if (h2d) for (size_t o = 0; o < IN_BYTES; o += chunk) CK(cudaMemcpyAsync(d_in + o, h_in + o, chunk, cudaMemcpyHostToDevice, s1));
                if (d2h) for (size_t o = 0; o < OUT_BYTES; o += chunk / 2) CK(cudaMemcpyAsync(h_out + o, d_out + o, chunk / 2, cudaMemcpyDeviceToHost, s2));
                CK(cudaStreamSynchronize(s1));
                CK(cudaStreamSynchronize(s2));
            };
            once();
            bar.wait();
            const double t0 = now_s();
            for (int r = 0; r < reps; r++) once();
            const double mine = now_s() - t0;
            bar.wait();
            const double all = now_s() - t0;
            per[g] = ((h2d ? IN_BYTES : 0) + (d2h ? OUT_BYTES : 0)) * (double)reps / mine / 1e9;
            if (g == 0) t_total = all;
            cudaStreamDestroy(s1); cudaStreamDestroy(s2);
            cudaFree(d_in); cudaFree(d_out); cudaFreeHost(h_in); cudaFreeHost(h_out);
        });
    }
    for (auto& t : th) t.join();
    const double bytes = ((h2d ? IN_BYTES : 0) + (d2h ? OUT_BYTES : 0)) * (double)reps * ngpus;
    printf("{\"test\": \"copy\", \"n_gpus\": %d, \"alloc\": \"%s\", \"dir\": \"%s\", \"chunk_MiB\": %zu, \"reps\": %d, "
           "\"ms_per_step\": %.3f, \"GBs_total\": %.1f, \"steps_per_s_total\": %.2f, \"per_gpu_GBs\": [",
           ngpus, alloc, h2d && d2h ? "both" : h2d ? "h2d" : "d2h", chunk / MB, reps, t_total / reps * 1e3, bytes / t_total / 1e9,
           ngpus * reps / t_total);
    for (int g = 0; g < ngpus; g++) printf("%s%.1f", g ? ", " : "", per[g]);
    printf("]}\n");
    fflush(stdout);
}

static void host_memcpy_bw(int threads) {
    const size_t bytes = 256 * MB;
    std::vector<char*> src(threads), dst(threads);
    for (int t = 0; t < threads; t++) { src[t] = (char*)malloc(bytes); dst[t] = (char*)malloc(bytes); memset(src[t], 1, bytes); memset(dst[t], 2, bytes); }
    Barrier bar(threads);
    double t_all = 0;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t]() {
            bar.wait();
            const double t0 = now_s();
            for (int r = 0; r < 4; r++) memcpy(dst[t], src[t], bytes);
            bar.wait();
            if (t == 0) t_all = now_s() - t0;
        });
    for (auto& x : th) x.join();
    printf("{\"test\": \"host_memcpy\", \"threads\": %d, \"GBs_copied\": %.1f, \"GBs_read_plus_write\": %.1f}\n", threads,
           4.0 * bytes * threads / t_all / 1e9, 8.0 * bytes * threads / t_all / 1e9);
    for (int t = 0; t < threads; t++) { free(src[t]); free(dst[t]); }
    fflush(stdout);
}

int main(int argc, char** argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    int maxg = argc > 1 ? atoi(argv[1]) : ndev;
    if (maxg > ndev) maxg = ndev;
    const bool quick = argc > 2 && !strcmp(argv[2], "quick"), full = argc > 2 && !strcmp(argv[2], "full");
    printf("{\"test\": \"topology\", \"gpus_visible\": %d, \"host_cpus_online\": %ld, \"numa_nodes_online\": \"%s\", \"gpus\": [", ndev,
           sysconf(_SC_NPROCESSORS_ONLN), read_file("/sys/devices/system/node/online").c_str());
    for (int g = 0; g < ndev; g++) {
        std::string pci;
        const int node = gpu_numa_node(g, &pci);
        cudaDeviceProp pr;
        CK(cudaGetDeviceProperties(&pr, g));
        printf("%s{\"index\": %d, \"pci\": \"%s\", \"numa_node\": %d, \"async_engines\": %d}", g ? ", " : "", g, pci.c_str(), node, pr.asyncEngineCount);
    }
    printf("]}\n");
    fflush(stdout);
    const int hw = (int)std::thread::hardware_concurrency();
    for (int t : {1, 4, 8, 16, 32})
        if (t <= hw) host_memcpy_bw(t);
    for (int n = 1; n <= maxg; n *= 2) {
        run_case(n, "default", true, true, 16 * MB, 8);   // the pipeline's chunking
        if (quick) continue;
        run_case(n, "default", true, true, 512 * MB, 8);  // one copy per operand pair
        run_case(n, "default", true, false, 512 * MB, 8);
        run_case(n, "default", false, true, 512 * MB, 8);
        if (!full) continue;
        run_case(n, "wc", true, true, 16 * MB, 8);
        run_case(n, "numa", true, true, 16 * MB, 8);
    }
    return 0;
}
