"""Named time markers with counter baselines (reference: utils/timings_tracker.py:22-74).

``BaseAgent.timings`` marks ``on_fit_start`` / ``on_train_epoch_start`` / ``on_validation_epoch_start`` with the collector's counters
like the reference (agents/base_agent.py:249-251, 297-299, 381-382); the reference's own ``DispatchMetricsCallback`` derives
``sys/timing/fps`` / ``fps_instant`` / ``eps`` from them through ``seconds_since`` / ``throughput_since``."""
from __future__ import annotations

import time
from typing import Any, Dict, Mapping, Optional, Tuple


def _numeric(values: Optional[Mapping[str, Any]]) -> Dict[str, float]:
    """int / float entries only (bools are ints, like the reference's isinstance check); arrays, None and strings are dropped."""
    return {k: float(v) for k, v in (values or {}).items() if isinstance(v, (int, float))}


class TimingsTracker:
    def __init__(self) -> None:
        self.markers: Dict[str, Tuple[int, Dict[str, float]]] = {}      # id -> (perf_counter_ns at start, counters at start)

    def start(self, marker_id: str, *, values: Optional[Mapping[str, Any]] = None) -> None:
        self.markers[marker_id] = (time.perf_counter_ns(), _numeric(values))

    def seconds_since(self, marker_id: str) -> float:
        """KeyError for a marker that was never started (fail fast, like the reference); never returns 0."""
        return max((time.perf_counter_ns() - self.markers[marker_id][0]) * 1e-9, 1e-12)

    def throughput_since(self, marker_id: str, *, values_now: Mapping[str, Any]) -> Dict[str, float]:
        """Per-second rate of every numeric counter in ``values_now`` since the marker (a counter absent at the start counts from 0;
        a counter that went down reports 0)."""
        base = self.markers[marker_id][1]
        dt = self.seconds_since(marker_id)
        return {k: max(v - base.get(k, 0.0), 0.0) / dt for k, v in _numeric(values_now).items()}
