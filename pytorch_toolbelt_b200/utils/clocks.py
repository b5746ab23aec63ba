"""SM clock / throttle-reason sampling DURING a timed region (bench.py's `clocks` key; B200_PROFILING.md recipe).

In-process NVML (nvidia_ml_py) polled every few milliseconds from a thread, so that even a 100 ms timed region gets
tens of samples; `nvidia-smi -lms` in a subprocess is the fallback when NVML cannot be opened.
"""
from __future__ import annotations

import statistics
import subprocess
import threading
import time

__all__ = ["ClockSampler"]

_REASON_BITS = {  # nvml.h nvmlClocksEventReason*
    "sw_power_cap": 0x4,
    "hw_slowdown": 0x8,
    "sw_thermal_slowdown": 0x20,
    "hw_thermal_slowdown": 0x40,
    "hw_power_brake_slowdown": 0x80,
}
_SMI_QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown")
_SMI_NAMES = ["sw_power_cap", "hw_slowdown", "sw_thermal_slowdown", "hw_thermal_slowdown"]


class ClockSampler:
    def __init__(self, index: int = 0, uuid: str | None = None, period_s: float = 0.004):
        self.index, self.uuid, self.period = index, uuid, period_s
        self.sm, self.max_mhz, self.reasons = [], None, set()
        self.source = "unavailable"
        self._stop = threading.Event()
        self._thread = None
        self._proc = None
        self._nvml = None
        self._handle = None

    # ---- NVML
    def _open_nvml(self) -> bool:
        try:
            import pynvml

            pynvml.nvmlInit()
            handle = None
            if self.uuid:
                try:
                    handle = pynvml.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
                except Exception:
                    handle = None
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
            self._nvml, self._handle = pynvml, handle
            self.source = "nvml"
            return True
        except Exception:
            return False

    def _poll_nvml(self):
        nv, h = self._nvml, self._handle
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    mask = int(get_reasons(h))
                    for name, bit in _REASON_BITS.items():
                        if mask & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    # ---- nvidia-smi fallback
    def _open_smi(self) -> bool:
        try:
            self._proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={_SMI_QUERY}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            return True
        except OSError:
            return False

    def _poll_smi(self):
        for line in self._proc.stdout:
            cells = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(float(cells[0]))
                self.max_mhz = max(self.max_mhz or 0.0, float(cells[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(_SMI_NAMES, cells[2:6]):
                if val.lower().startswith("active"):
                    self.reasons.add(name)

    def __enter__(self):
        if self._open_nvml():
            self._thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self._thread.start()
        elif self._open_smi():
            self._thread = threading.Thread(target=self._poll_smi, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        if self._proc is not None:
            time.sleep(0.12)  # let one more -lms period land
            self._proc.terminate()
            try:
                self._proc.wait(timeout=2)
            except Exception:
                self._proc.kill()
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if self._nvml is not None:
            try:
                self._nvml.nvmlShutdown()
            except Exception:
                pass

    @property
    def samples(self) -> int:
        return len(self.sm)

    def summary(self) -> dict:
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0, "source": self.source}
        return {"sm_mhz": float(statistics.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": self.source}
