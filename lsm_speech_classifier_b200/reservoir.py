"""Reservoir definition: parameters and the host-side builder.

The reference constructs its liquid through the un-vendored ``snnpy`` package:
``SimulationParams(...)`` at /root/reference/extract_lsm_features.py:164-175, two attribute
writes at :185-186 and ``SNN(simulation_params=...)`` at :188.  The reservoir is built ONCE
and reused for every utterance (:188 is outside the loops at :78), so it is plain read-only
data for the GPU.  We build it on the host with one ``numpy.random.RandomState`` consumed in
a documented order and upload the arrays; the CPU oracle is handed the very same arrays.

Frozen spec (DESIGN.md "Reservoir spec", R1-R5,R7; snnpy parity is UNPINNED):

R1  one RandomState(seed) consumed in this order: topology, weights, input map, output set, leaks.
R2  Watts-Strogatz small world over N neurons: ring lattice with k/2 neighbours each side,
    each "right-hand" edge (u, u+j) rewired with probability p to a uniformly random
    non-neighbour; undirected pattern used in both directions, no self loops.
R3  one weight per DIRECTED edge, drawn in (postsynaptic, presynaptic) row-major order from
    Normal(mean_weight, |mean_weight| / weight_variance), then rounded to the nearest multiple
    of 2**-24 and stored as int32 (Q7.24).  Row sums of such numbers are exact in fp64 (and in
    int32/int64) whatever the summation order, so every correct implementation - serial CPU,
    event-driven GPU, tensor-core GPU - produces the same bits.
R4  input row r (C*R rows) drives exactly one reservoir neuron, distinct rows -> distinct
    neurons while C*R <= N, with current `input_gain` (default: the membrane threshold, i.e. an
    active input step makes its neuron fire unless refractory) per active time step (level signal).
R5  N_out output neurons chosen without replacement, kept ascending.
R7  leak_i = leak_coefficient, or Normal(leak, leak/d) clipped to [0,1] when a
    leak_variance_divisor d is given.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

W_SHIFT = 24  # weights are int32 multiples of 2**-24


@dataclass
class SimulationParams:
    """Field names follow the keyword arguments the reference passes
    (extract_lsm_features.py:164-175) plus the two it sets afterwards (:185-186)."""
    num_neurons: int = 1000
    mean_weight: float = 0.0
    num_output_neurons: int = 400
    membrane_threshold: float = 2.0
    leak_coefficient: float = 1 / 100
    refractory_period: int = 2
    small_world_graph_p: float = 0.1
    small_world_graph_k: int = 200
    input_spike_times: Optional[np.ndarray] = None   # uint8[C*R, T]: only its shape is used at build time
    leak_variance_divisor: Optional[float] = None
    weight_variance: float = 10.0
    # not in the reference's call: ours, with defaults that keep its behaviour
    input_gain: Optional[float] = None   # None -> membrane_threshold: an active input step makes its neuron fire
    seed: int = 42
    quantize_weights: bool = True        # R3: weights rounded to 2**-24 (int32, order-free sums).  False: SURVEY.md 8c S3/S6 as
                                         # written - fp64 normals, row sums in ascending presynaptic order ("strict" reservoirs)


@dataclass
class ReservoirDef:
    """Plain arrays; everything the oracle and the CUDA library need."""
    num_neurons: int
    num_inputs: int
    num_steps: int
    theta: float
    refractory: int
    w_rowptr: np.ndarray   # int32[N+1]   CSR over postsynaptic neuron
    w_col: np.ndarray      # int32[nnz]   presynaptic neuron, ascending inside a row
    w_q: np.ndarray        # int32[nnz]   weight * 2**24
    in_rowptr: np.ndarray  # int32[N+1]   CSR over reservoir neuron
    in_col: np.ndarray     # int32[nin]   input row, ascending inside a row
    in_val: np.ndarray     # float64[nin]
    out_idx: np.ndarray    # int32[N_out] ascending
    leak: np.ndarray       # float64[N]
    w_val: Optional[np.ndarray] = None   # float64[nnz]: strict reservoirs (quantize_weights=False) carry their weights here, w_q is zeros
    w_shift: int = W_SHIFT
    meta: dict = field(default_factory=dict)


def watts_strogatz_adjacency(n: int, k: int, p: float, rs: np.random.RandomState) -> np.ndarray:
    """bool[n,n] symmetric adjacency, zero diagonal (R2)."""
    half = k // 2
    if half >= n / 2:
        raise ValueError("small_world_graph_k must be < num_neurons")
    adj = np.zeros((n, n), dtype=bool)
    idx = np.arange(n)
    for j in range(1, half + 1):
        adj[idx, (idx + j) % n] = True
        adj[(idx + j) % n, idx] = True
    if p > 0:
        deg = adj.sum(axis=1)
        for j in range(1, half + 1):
            draw = rs.random_sample(n)
            for u in np.nonzero(draw < p)[0]:
                v = (u + j) % n
                if deg[u] >= n - 1 or not adj[u, v]:
                    continue
                while True:
                    w = rs.randint(n)
                    if w != u and not adj[u, w]:
                        break
                adj[u, v] = adj[v, u] = False
                adj[u, w] = adj[w, u] = True
                deg[v] -= 1
                deg[w] += 1
    return adj


def build_reservoir(p: SimulationParams) -> ReservoirDef:
    if p.input_spike_times is None:
        raise ValueError("input_spike_times (one sample, uint8[C,T]) is required to size the input map")
    n = int(p.num_neurons)
    rows, steps = (int(v) for v in np.asarray(p.input_spike_times).shape)
    rs = np.random.RandomState(p.seed)                                    # R1
    adj = watts_strogatz_adjacency(n, int(p.small_world_graph_k), float(p.small_world_graph_p), rs)  # R2
    post, pre = np.nonzero(adj)                                           # row-major: post asc, pre asc
    nnz = len(post)
    sd = abs(p.mean_weight) / p.weight_variance if p.weight_variance else 0.0
    w = rs.normal(p.mean_weight, sd, size=nnz) if sd > 0 else np.full(nnz, float(p.mean_weight))
    if p.quantize_weights:
        wq = np.rint(w * float(1 << W_SHIFT))
        if np.any(np.abs(wq) >= 2 ** 31):
            raise ValueError("weight out of Q7.24 range")
        wq = wq.astype(np.int32)                                          # R3
        w_val = None
    else:
        wq = np.zeros(nnz, dtype=np.int32)                                # S3: the fp64 draws themselves
        w_val = np.ascontiguousarray(w, dtype=np.float64)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(post, minlength=n), out=rowptr[1:])
    # worst-case row sum must stay inside int32 so any integer accumulator is exact
    absrow = np.add.reduceat(np.abs(wq.astype(np.int64)), rowptr[:-1][rowptr[:-1] < nnz]) if nnz else np.zeros(1)
    if absrow.max(initial=0) >= 2 ** 31:
        raise ValueError("row sum of |weights| exceeds the exact int32 range")
    if rows <= n:
        in_neuron = rs.permutation(n)[:rows]                              # R4
    else:
        in_neuron = rs.randint(n, size=rows)
    order = np.lexsort((np.arange(rows), in_neuron))
    in_col = np.arange(rows, dtype=np.int32)[order]
    in_rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(in_neuron, minlength=n), out=in_rowptr[1:])
    gain = float(p.membrane_threshold if p.input_gain is None else p.input_gain)
    in_val = np.full(rows, gain)
    n_out = min(int(p.num_output_neurons), n)
    out_idx = np.sort(rs.permutation(n)[:n_out]).astype(np.int32)         # R5
    if p.leak_variance_divisor:
        leak = rs.normal(p.leak_coefficient, p.leak_coefficient / p.leak_variance_divisor, size=n)
        leak = np.clip(leak, 0.0, 1.0)                                    # R7
    else:
        leak = np.full(n, float(p.leak_coefficient))
    return ReservoirDef(
        num_neurons=n, num_inputs=rows, num_steps=steps,
        theta=float(p.membrane_threshold), refractory=int(p.refractory_period),
        w_rowptr=rowptr, w_col=pre.astype(np.int32), w_q=wq,
        in_rowptr=in_rowptr, in_col=in_col, in_val=in_val, out_idx=out_idx,
        leak=np.ascontiguousarray(leak, dtype=np.float64), w_val=w_val,
        meta=dict(seed=p.seed, k=int(p.small_world_graph_k), p=float(p.small_world_graph_p),
                  mean_weight=float(p.mean_weight), weight_variance=float(p.weight_variance),
                  input_gain=gain),
    )
