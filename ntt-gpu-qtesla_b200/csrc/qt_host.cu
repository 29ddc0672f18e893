// qt_host.cu — host-side runtime helpers of the C ABI that launch no kernel: where a GPU hangs in the
// machine (PCI bus id, NUMA node from sysfs) and binding the calling host thread next to it, so that
// the pinned buffers a caller allocates afterwards — and the copy threads of qt_polymul_host_multi — live
// on the memory controller the GPU's PCIe root hangs off.  (The reference has no counterpart: one GPU,
// pageable malloc buffers, synchronous cudaMemcpy, NTT.cu:2105-2124.)
#include <cuda_runtime.h>
#include <sched.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/qtesla_b200.h"

namespace {

std::string slurp(const std::string& path) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return "";
    char buf[8192];
    size_t k = fread(buf, 1, sizeof(buf) - 1, f);
    fclose(f);
    while (k && (buf[k - 1] == '\n' || buf[k - 1] == ' ')) k--;
    buf[k] = 0;
    return buf;
}

// "0-15,32-47" -> cpu_set_t; returns the number of CPUs
int parse_cpulist(const std::string& s, cpu_set_t* set) {
    CPU_ZERO(set);
    int count = 0;
    const char* p = s.c_str();
    while (*p) {
        char* e;
        long a = strtol(p, &e, 10), b = a;
        if (e == p) break;
        if (*e == '-') {
            p = e + 1;
            b = strtol(p, &e, 10);
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) {
            CPU_SET((int)c, set);
            count++;
        }
        if (*e != ',') break;
        p = e + 1;
    }
    return count;
}

}  // namespace

extern "C" {

int qt_device_pci_bus_id(int device, char* out, size_t len) {
    if (!out || len < 13) return QT_ERR_BAD_ARG;
    cudaError_t e = cudaDeviceGetPCIBusId(out, (int)len, device);
    if (e != cudaSuccess) return (int)e;
    for (char* c = out; *c; c++) *c = (char)tolower(*c);  // sysfs spells it in lower case
    return 0;
}

int qt_device_numa_node(int device, int* node_out) {
    if (!node_out) return QT_ERR_BAD_ARG;
    char pci[32];
    int rc = qt_device_pci_bus_id(device, pci, sizeof(pci));
    if (rc) return rc;
    const std::string s = slurp(std::string("/sys/bus/pci/devices/") + pci + "/numa_node");
    *node_out = s.empty() ? -1 : atoi(s.c_str());  // -1: the platform (or the hypervisor) does not say
    return 0;
}

int qt_bind_thread_to_device(int device, int* cpus_out) {
    if (cpus_out) *cpus_out = 0;
    int node = -1;
    int rc = qt_device_numa_node(device, &node);
    if (rc) return rc;
    if (node < 0) return 0;  // unknown: leave the thread where it is
    cpu_set_t want, have, both;
    if (!parse_cpulist(slurp("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist"), &want)) return 0;
    if (sched_getaffinity(0, sizeof(have), &have) != 0) return 0;
    CPU_AND(&both, &want, &have);  // never leave the cpuset the process was given (containers, taskset)
    const int n = CPU_COUNT(&both);
    if (n == 0) return 0;
    if (sched_setaffinity(0, sizeof(both), &both) != 0) return 0;
    if (cpus_out) *cpus_out = n;
    return 0;
}

}  // extern "C"
