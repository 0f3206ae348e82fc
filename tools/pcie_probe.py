"""Measure pinned H2D / D2H bandwidth per NUMA node of the host buffer (dev tool, not part of the product)."""
import glob, os, time, subprocess
import torch

def cpus_of(node):
    s = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in s.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out

nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes", nodes, "cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
for f in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    try:
        cls = open(os.path.dirname(f) + "/class").read().strip()
        if cls.startswith("0x0302") or cls.startswith("0x0300"):
            print(f, open(f).read().strip())
    except Exception:
        pass
torch.cuda.init()
N = 1 << 30
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
dev2 = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
all_cpus = os.sched_getaffinity(0)
for node in nodes:
    cp = [c for c in cpus_of(node) if c in all_cpus]
    if not cp:
        print("node", node, "no allowed cpus"); continue
    os.sched_setaffinity(0, cp)
    h = torch.empty(N, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(N, dtype=torch.uint8).pin_memory()
    h.fill_(1); h2.fill_(2)
    for name, fn in (("h2d", lambda: dev.copy_(h, non_blocking=True)),
                     ("d2h", lambda: h2.copy_(dev2, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize()
        print(f"node {node} {name}: {5 * N / (time.perf_counter() - t) / 1e9:.1f} GB/s")
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1): dev.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(dev2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"node {node} bidir: {5 * N / dt / 1e9:.1f} GB/s each direction")
    del h, h2
    os.sched_setaffinity(0, all_cpus)
