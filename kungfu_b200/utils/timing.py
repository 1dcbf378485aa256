"""Measurement helpers: device timing (CUDA events, max over ranks) and clock/throttle sampling.

The reference times with host wall clocks (`timeit`, benchmarks/system/benchmark_kungfu_torch.py:118-127); every
number we report is device-timed and, for multi-GPU runs, the max over ranks.
"""
from __future__ import annotations

import shutil
import statistics
import subprocess
import threading
from typing import Dict, List, Optional

import torch

_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
          "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
          "clocks_event_reasons.sw_power_cap")


class ClockSampler:
    """Samples `nvidia-smi` clocks and throttle reasons for one GPU while a timed region runs."""

    def __init__(self, gpu_index: int = 0, period_ms: int = 200) -> None:
        self.gpu_index = gpu_index
        self.period_ms = period_ms
        self.rows: List[List[str]] = []
        self._proc: Optional[subprocess.Popen] = None
        self._thread: Optional[threading.Thread] = None

    def start(self) -> "ClockSampler":
        exe = shutil.which("nvidia-smi")
        if exe is None:
            return self
        try:
            self._proc = subprocess.Popen(
                [exe, f"--query-gpu={_QUERY}", "--format=csv,noheader,nounits", "-i", str(self.gpu_index), "-lms",
                 str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self._proc = None
            return self

        def pump() -> None:
            assert self._proc is not None and self._proc.stdout is not None
            for line in self._proc.stdout:
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 9:
                    self.rows.append(parts)

        self._thread = threading.Thread(target=pump, daemon=True)
        self._thread.start()
        return self

    def stop(self) -> Dict:
        if self._proc is not None:
            try:
                self._proc.terminate()  # exact PID we started
                self._proc.wait(timeout=5)
            except Exception:
                try:
                    self._proc.kill()
                except Exception:
                    pass
        if self._thread is not None:
            self._thread.join(timeout=2)
        return self.summary()

    def summary(self) -> Dict:
        sm, smax, power = [], [], []
        reasons = set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "reasons": sorted(reasons), "samples": len(sm)}


def max_over_ranks(value: float) -> float:
    """Max of a host scalar over all ranks (control plane)."""
    from ..runtime import get_runtime

    rt = get_runtime()
    if rt.size == 1:
        return value
    import struct

    parts = rt.control.all_gather_bytes(struct.pack("<d", float(value)))
    return max(struct.unpack("<d", p)[0] for p in parts)


class DeviceTimer:
    """CUDA-event stopwatch on the current stream."""

    def __init__(self) -> None:
        self.t0 = torch.cuda.Event(enable_timing=True)
        self.t1 = torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        torch.cuda.synchronize()
        self.t0.record()
        return self

    def __exit__(self, *exc):
        self.t1.record()
        torch.cuda.synchronize()
        self.ms = self.t0.elapsed_time(self.t1)
        return False


def flush_l2(buf: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Write a buffer larger than L2 (126 MB) so the next timed iteration starts cold."""
    if buf is None:
        buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    buf.fill_(1)
    return buf


def nvlink_counters(gpu_index: int = 0) -> Optional[Dict[str, float]]:
    """Cumulative NVLink data counters of one GPU, summed over its links, in bytes: {"tx": ..., "rx": ...} from
    `nvidia-smi nvlink -gt d` (None if the tool or the counters are unavailable). Read them before and after a loop of K
    launches: (after - before) / K is what one launch really moved over the links — the hardware's own byte count next to the
    algorithmic one (Nsight Compute cannot attach to kernels that touch peer-mapped / multicast memory: `ncu` aborts with
    "UnknownError" on them, see profiles/collectives_ncu.md)."""
    import re

    exe = shutil.which("nvidia-smi")
    if exe is None:
        return None
    try:
        out = subprocess.run([exe, "nvlink", "-gt", "d", "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout
    except Exception:
        return None
    tx = rx = 0.0
    seen = False
    for line in out.splitlines():
        m = re.search(r"Data\s+(Tx|Rx):\s*([0-9]+)\s*(KiB|MiB|GiB|B)?", line, re.I)
        if not m:
            continue
        mult = {"b": 1.0, "kib": 1024.0, "mib": 1048576.0, "gib": 1073741824.0}.get((m.group(3) or "KiB").lower(), 1024.0)
        v = float(m.group(2)) * mult
        seen = True
        if m.group(1).lower() == "tx":
            tx += v
        else:
            rx += v
    return {"tx": tx, "rx": rx} if seen else None
