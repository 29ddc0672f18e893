"""Host placement helpers over the C ABI (qt_device_pci_bus_id / qt_device_numa_node /
qt_bind_thread_to_device, csrc/qt_host.cu): a rank that feeds one GPU from host buffers binds itself to the
CPUs of that GPU's NUMA node BEFORE it allocates its pinned arrays, so first-touch places them next to the
GPU's PCIe root.  Unknown topology (a VM that reports node -1) leaves the process alone."""
import ctypes as C

from .engine import lib, _check


def gpu_pci_bus_id(device):
    buf = C.create_string_buffer(32)
    _check(lib().qt_device_pci_bus_id(device, buf, 32))
    return buf.value.decode()


def gpu_numa_node(device):
    node = C.c_int(-1)
    _check(lib().qt_device_numa_node(device, C.byref(node)))
    return node.value


def bind_to_gpu_node(device):
    """Binds the calling thread (the main thread of a one-process-per-GPU rank) to the GPU's NUMA node.
    Returns the number of CPUs bound to, 0 when nothing was changed."""
    cpus = C.c_int(0)
    _check(lib().qt_bind_thread_to_device(device, C.byref(cpus)))
    return cpus.value
