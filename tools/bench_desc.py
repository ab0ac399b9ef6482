#!/usr/bin/env python
"""Micro-benchmark of K-b (f4l_desc_nn) on one GPU: tensor-core path TFLOP/s (2*N*M*D per direction)
against MEASURED_PEAKS.json, kernel times from the library's in-stream events.
    python tools/bench_desc.py [--n 262144] [--m 262144] [--d 64] [--reps 3]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion4landslide_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=262144)
    ap.add_argument("--m", type=int, default=262144)
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--algo", default="tensor")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    A = torch.nn.functional.normalize(torch.randn(a.n, a.d, device=dev, generator=g), dim=1)
    B = torch.nn.functional.normalize(torch.randn(a.m, a.d, device=dev, generator=g), dim=1)
    k = min(a.n, a.m) // 2
    A[:k] = torch.nn.functional.normalize(B[:k] + 0.15 * torch.randn(k, a.d, device=dev, generator=g), dim=1)
    ops.desc_nn(A, B, algo=a.algo)
    torch.cuda.synchronize()
    L = _lib.lib()
    L.f4l_profile_reset()
    L.f4l_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        idx, d2 = ops.desc_nn(A, B, algo=a.algo)
    e1.record()
    torch.cuda.synchronize()
    L.f4l_profile_enable(0)
    tab = _lib.profile_table()
    ms = e0.elapsed_time(e1) / a.reps
    flop = 2.0 * a.n * a.m * a.d
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    out = {"n": a.n, "m": a.m, "d": a.d, "algo": a.algo, "ms_per_call": ms, "tflops_call": flop / ms / 1e9,
           "kernels": {k: {"ms_avg": v[0] / v[1], "launches": v[1]} for k, v in tab.items()},
           "matched_to_first_half": float((idx[:k] == torch.arange(k, device=dev)).float().mean())}
    if "k_desc_nn_tc" in tab:
        t = tab["k_desc_nn_tc"][0] / tab["k_desc_nn_tc"][1]
        out["tc_kernel_tflops"] = flop / t / 1e9
        if peaks:
            out["frac_of_measured_bf16_burst"] = out["tc_kernel_tflops"] / peaks["bf16_tflops"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
