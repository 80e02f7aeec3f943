"""Pinned-memory copy bandwidth of this box, per NUMA node the host buffers are allocated from (python tools/pcie_duplex.py):
H2D alone, D2H alone, both at once (what bench.py's e2e leg does every step).  The process binds itself to the CPUs of one
node before it allocates (first touch decides where pinned pages live)."""
import glob
import os
import subprocess
import sys

import torch


def cpus_of(node):
    out = []
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def measure(tag, nbytes=512 << 20, reps=6):
    dev = torch.device("cuda", 0)
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); h_out.fill_(2)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.ones(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        cur = torch.cuda.current_stream(dev)
        cur.wait_stream(s1); cur.wait_stream(s2)
        e1.record(); torch.cuda.synchronize()
        return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    run(True, True)
    print(f"{tag}: H2D alone {run(True, False):.1f} GB/s, D2H alone {run(False, True):.1f} GB/s, both at once {run(True, True):.1f} GB/s per direction", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        node = int(sys.argv[1])
        os.sched_setaffinity(0, cpus_of(node))
        measure(f"host buffers on NUMA node {node}")
    else:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
        nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
        print("NUMA nodes:", nodes, "cpus:", {n: len(cpus_of(n)) for n in nodes}, "this process may run on", len(os.sched_getaffinity(0)), "cpus", flush=True)
        measure("no binding")
        for n in nodes:
            subprocess.run([sys.executable, __file__, str(n)])
