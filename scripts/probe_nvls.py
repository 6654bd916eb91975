#!/usr/bin/env python
"""Is an NVLS (NVLink SHARP multicast) all-gather an option on this box?  Prints the driver's multicast attribute per
GPU and whether NCCL itself brings NVLS up (NCCL_DEBUG=INFO under torchrun).  python scripts/probe_nvls.py"""
import os
import subprocess
import sys


def driver_attribute():
    try:
        from cuda.bindings import driver as drv
    except ImportError:
        from cuda import cuda as drv
    (err,) = drv.cuInit(0)
    err, n = drv.cuDeviceGetCount()
    out = []
    for i in range(n):
        err, dev = drv.cuDeviceGet(i)
        attr = drv.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED
        err, v = drv.cuDeviceGetAttribute(attr, dev)
        attr2 = drv.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED
        err, f = drv.cuDeviceGetAttribute(attr2, dev)
        out.append((i, int(v), int(f)))
    return out


CHILD = r'''
import os, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(1 << 24, device="cuda"); out = torch.empty(x.numel() * dist.get_world_size(), device="cuda")
for _ in range(3):
    dist.all_gather_into_tensor(out, x)
torch.cuda.synchronize(); dist.destroy_process_group()
'''

if __name__ == "__main__":
    print("multicast / fabric-handle support per GPU (index, multicast, fabric):", driver_attribute(), flush=True)
    import torch

    n = torch.cuda.device_count()
    if n >= 2:
        import tempfile

        env = dict(os.environ, NCCL_DEBUG="INFO", NCCL_DEBUG_SUBSYS="INIT,NVLS")
        child = os.path.join(tempfile.mkdtemp(prefix="nvls_"), "child.py")
        open(child, "w").write(CHILD)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                            "--master-addr", "127.0.0.1", "--master-port", "29533", child],
                           env=env, capture_output=True, text=True, timeout=300)
        lines = [ln for ln in (r.stdout + r.stderr).splitlines() if "NVLS" in ln or "nvls" in ln]
        print(f"NCCL lines mentioning NVLS ({len(lines)}):")
        for ln in lines[:12]:
            print("  ", ln[:200])
        if not lines:
            print("   none (tail of the log follows)")
            print("\n".join((r.stdout + r.stderr).splitlines()[-8:]))
