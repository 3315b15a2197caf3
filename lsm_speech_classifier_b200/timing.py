"""Wall-clock breakdown of a pipeline run (main.py --timing): where the time goes between audio and the feature file."""
from __future__ import annotations

import time
from contextlib import contextmanager

enabled = False
_stages: dict[str, float] = {}
_units: dict[str, int] = {}


@contextmanager
def stage(name: str, utterances: int | None = None):
    t0 = time.perf_counter()
    try:
        yield
    finally:
        if enabled:
            _stages[name] = _stages.get(name, 0.0) + time.perf_counter() - t0
            if utterances is not None:
                _units[name] = _units.get(name, 0) + int(utterances)


def report() -> str:
    total = sum(_stages.values())
    lines = ["| stage | seconds | share | utterances/s |", "|---|---|---|---|"]
    for k, v in _stages.items():
        rate = f"{_units[k] / v:,.0f}" if k in _units and v > 0 else ""
        lines.append(f"| {k} | {v:.3f} | {100 * v / max(total, 1e-12):.1f} % | {rate} |")
    lines.append(f"| total | {total:.3f} | | |")
    return "\n".join(lines)
