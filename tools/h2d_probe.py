"""Host -> device feed of a multi-GPU box: what limits the e2e number at N = 4 and 8.

  1. topology: NUMA nodes, CPUs and memory per node, the NUMA node and PCIe link of every GPU;
  2. H2D rate of every GPU alone (pinned buffer allocated by cudaHostAlloc from this thread);
  3. all GPUs at once, same buffers;
  4. the same two with the buffer of GPU g placed on NUMA node k by set_mempolicy(MPOL_BIND) + first touch +
     cudaHostRegister, for every k: shows whether placement matters and which node feeds which GPU best;
  5. write-combined pinned buffers (cudaHostAllocWriteCombined), all GPUs at once.

usage: python tools/h2d_probe.py [MB per buffer]      (one process, one stream per device; CUDA events)"""
import ctypes
import glob
import os
import sys
import time

import numpy as np
import torch

MB = int(sys.argv[1]) if len(sys.argv) > 1 else 512
NB = MB << 20
rt = torch.cuda.cudart()
libc = ctypes.CDLL(None, use_errno=True)
ng = torch.cuda.device_count()


def sh(cmd):
    return os.popen(cmd + " 2>&1").read().strip()


print("== topology")
nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
for k in nodes:
    cpus = open("/sys/devices/system/node/node%d/cpulist" % k).read().strip()
    mem = [l for l in open("/sys/devices/system/node/node%d/meminfo" % k) if "MemTotal" in l or "MemFree" in l]
    print("node %d: cpus %s; %s" % (k, cpus, "; ".join(" ".join(l.split()[2:]) for l in mem)))
print("this process may run on cpus:", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)), "cpus")
for g in range(ng):
    bdf = torch.cuda.get_device_properties(g).pci_bus_id if hasattr(torch.cuda.get_device_properties(g), "pci_bus_id") else None
q = sh("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader")
for line in q.splitlines():
    idx, bdf = [x.strip() for x in line.split(",")[:2]]
    p = "/sys/bus/pci/devices/%s/numa_node" % bdf.lower().replace("00000000:", "0000:")
    print("gpu", line, "| numa_node", open(p).read().strip() if os.path.exists(p) else "?")
print(sh("nvidia-smi topo -m | head -14"))

streams = [torch.cuda.Stream(device=g) for g in range(ng)]
dev = [torch.empty(NB, dtype=torch.uint8, device="cuda:%d" % g) for g in range(ng)]


def rate(hosts, gpus, reps=6):
    """GB/s per GPU with the copies of all `gpus` in flight together."""
    evs = {}
    for g in gpus:
        with torch.cuda.device(g):
            dev[g].copy_(hosts[g], non_blocking=True)       # warm
    for g in gpus:
        torch.cuda.synchronize(g)
    for g in gpus:
        with torch.cuda.device(g), torch.cuda.stream(streams[g]):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                dev[g].copy_(hosts[g], non_blocking=True)
            e1.record()
            evs[g] = (e0, e1)
    out = {}
    for g in gpus:
        torch.cuda.synchronize(g)
        out[g] = NB * reps / evs[g][0].elapsed_time(evs[g][1]) / 1e6
    return out


def fmt(r):
    return " ".join("%5.1f" % r[g] for g in sorted(r)) + "  | sum %6.1f GB/s" % sum(r.values())


print("\n== pinned by cudaHostAlloc from this thread (%d MB per GPU)" % MB)
pinned = [torch.empty(NB, dtype=torch.uint8).pin_memory() for _ in range(ng)]
print("alone      :", fmt({g: rate(pinned, [g])[g] for g in range(ng)}))
for grp in ([0, 1], [0, 1, 2, 3], [4, 5, 6, 7], list(range(ng))):
    grp = [g for g in grp if g < ng]
    if len(grp) > 1:
        print("together %-2d:" % len(grp), fmt(rate(pinned, grp)))
del pinned

MPOL_DEFAULT, MPOL_BIND = 0, 2


def set_mempolicy(mode, node=None):
    if node is None:
        rc = libc.syscall(238, mode, None, 0)
    else:
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(238, mode, ctypes.byref(mask), 64)
    return rc


def registered_on(node):
    if set_mempolicy(MPOL_BIND, node) != 0:
        return None
    a = np.empty(NB, np.uint8); a[::4096] = 1          # first touch under the policy
    set_mempolicy(MPOL_DEFAULT)
    t = torch.from_numpy(a)
    rc = rt.cudaHostRegister(t.data_ptr(), NB, 0)
    return t if int(rc) == 0 else None


if len(nodes) > 1:
    print("\n== buffer bound to NUMA node k (set_mempolicy + first touch + cudaHostRegister): GB/s of GPU g alone")
    best = {}
    for k in nodes:
        bufs = registered_on(k)
        if bufs is None:
            print("node %d: set_mempolicy / cudaHostRegister not permitted here" % k); continue
        r = {g: rate({g: bufs}, [g])[g] for g in range(ng)}
        print("node %d    :" % k, fmt(r))
        for g in r:
            if r[g] > best.get(g, (0, 0))[0]:
                best[g] = (r[g], k)
        rt.cudaHostUnregister(bufs.data_ptr())
    if best:
        print("best node per GPU:", {g: best[g][1] for g in sorted(best)})
        hosts = {g: registered_on(best[g][1]) for g in range(ng)}
        if all(v is not None for v in hosts.values()):
            print("all together, every buffer on its GPU's best node:", fmt(rate(hosts, list(range(ng)))))
else:
    print("\n(one NUMA node: placement cannot matter)")

print("\n== write-combined pinned buffers (cudaHostAllocWriteCombined)")
wc = []
for g in range(ng):
    p = ctypes.c_void_p()
    rc = ctypes.CDLL("libcudart.so.12" if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else "libcudart.so").cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(NB), ctypes.c_uint(4))
    if rc != 0:
        wc = None; break
    ctypes.memset(p, 1, NB)
    wc.append(torch.frombuffer((ctypes.c_uint8 * NB).from_address(p.value), dtype=torch.uint8))
if wc:
    print("alone      :", fmt({g: rate(wc, [g])[g] for g in range(ng)}))
    print("together %-2d:" % ng, fmt(rate(wc, list(range(ng)))))
