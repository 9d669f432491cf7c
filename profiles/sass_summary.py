#!/usr/bin/env python
"""Mnemonic counts per kernel of the shipped library:  python profiles/sass_summary.py > profiles/r2/sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "wannierberri_b200", "libwbgpu.so")], capture_output=True, text=True).stdout
print("# cuobjdump -sass wannierberri_b200/libwbgpu.so (sm_100a cubin), mnemonic counts per kernel")
print("# DMMA = FP64 tensor-core mma.sync.m8n8k4; UBLKCP = TMA bulk copy (cp.async.bulk); SYNCS = mbarrier ops;")
print("# DFMA/DMUL/DADD = FP64 vector pipe; MUFU.RSQ64H / RCP64H = seeds of rsqrt / __drcp_rn")
print("# (tcgen05 has no FP64 kind: UTC*MMA / LDTM cannot appear in an FP64 path)")
cols = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", r"MUFU\.RSQ64H", r"MUFU\.RCP64H", "SHFL", "LDS", "STS", "LDG", "STG", "BAR"]
print(f"# {'kernel':70s} {'instr':>7s} " + " ".join(f"{c.replace(chr(92), '').replace('MUFU.', ''):>6s}" for c in cols))
funcs = re.split(r"\n\s*Function : ", txt)[1:]
keep = ("wb_events_mma_kernel<18", "wb_axis10_fused_kernel<20, 128", "wb_axis_dft_kernel<10>", "wb_tridiag2", "wb_trideig_kernel<18",
        "wb_backtransform_kernel<18", "wb_rotate_gemm", "wb_kubo_accumulate_optcond_tiled", "wb_tridiag_cta", "wb_eigvec_cta",
        "wb_scan_accumulate", "wb_tridiag_tpm2_kernel<18", "wb_events_mma_kernel<24", "wb_trideig_kernel<24", "wb_events_xbar",
        "wb_deromega", "wb_dermorb", "wb_product_events", "wb_tetra_accumulate", "wb_kubo_entries")
tot = collections.Counter()
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
    n = len(re.findall(r"/\*[0-9a-f]{4,6}\*/\s+[A-Z@]", f))
    c = [len(re.findall(r"\b" + p + r"\b", f)) for p in cols]
    tot.update({"kernels": 1, "DMMA": c[0], "UBLKCP": c[4], "SYNCS": c[5]})
    if any(k in d for k in keep):
        print(f"{d[:70]:72s} {n:7d} " + " ".join(f"{x:6d}" for x in c))
print("# architectures in the binary:", sorted(set(re.findall(r"arch = (sm_\w+)", txt))))
print("# whole library:", dict(tot))
