"""PCIe copy rates on this box: H2D alone, D2H alone, both at once, from pinned memory allocated (a) wherever the process
happens to run and (b) after binding the process to the CPUs of the GPU's NUMA node.  Diagnostics for the e2e pipeline."""
import glob
import os
import sys
import time

import torch


def gpu_numa_cpus(dev=0):
    bdf = torch.cuda.get_device_properties(dev).pci_bus_id if hasattr(torch.cuda.get_device_properties(dev), "pci_bus_id") else None
    try:
        import subprocess
        out = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(dev)], text=True).strip()
        bdf = out.lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        cpus = open(f"/sys/devices/system/node/node{max(node, 0)}/cpulist").read().strip()
        return node, cpus
    except Exception as ex:
        return None, str(ex)


def parse_cpulist(s):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def measure(tag, nbytes=8 << 30, reps=3):
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    h_out.fill_(2)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return nbytes * reps / (time.perf_counter() - t0) / 1e9

    run(True, True)
    a, b, c = run(True, False), run(False, True), run(True, True)
    print(f"{tag}: H2D alone {a:.1f} GB/s, D2H alone {b:.1f} GB/s, both at once {c:.1f} GB/s per direction", flush=True)
    del h_in, h_out, d_in, d_out


if __name__ == "__main__":
    torch.cuda.init()
    print("cpus allowed:", len(os.sched_getaffinity(0)), "numa nodes:", len(glob.glob("/sys/devices/system/node/node[0-9]*")), flush=True)
    node, cpus = gpu_numa_cpus(0)
    print("gpu0 numa node:", node, "cpus:", cpus, flush=True)
    measure("default placement")
    if node is not None and node >= 0:
        os.sched_setaffinity(0, set(parse_cpulist(cpus)) & os.sched_getaffinity(0))
        measure(f"bound to node {node}")
        other = [n for n in range(len(glob.glob('/sys/devices/system/node/node[0-9]*'))) if n != node]
        if other:
            oc = open(f"/sys/devices/system/node/node{other[-1]}/cpulist").read().strip()
            os.sched_setaffinity(0, set(parse_cpulist(oc)))
            measure(f"bound to node {other[-1]} (remote)")
