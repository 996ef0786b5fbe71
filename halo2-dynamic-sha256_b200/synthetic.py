"""Counter-based synthetic messages: message m of a workload is a pure function of (seed, m), so the
host, the oracle and every GPU shard derive identical bytes and lengths with no transfer.

    len(m)   = lo + splitmix64(seed ^ LEN_TAG + m) % (hi - lo + 1)
    bytes(m) = little-endian bytes of splitmix64((seed + m * GOLDEN) + j) for j = 0, 1, ...
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

SEED = 0x5348413235360001
GOLDEN = 0x9E3779B97F4A7C15
LEN_TAG = 0xA5A5A5A55A5A5A5A
M64 = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised SplitMix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(GOLDEN)).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json configuration."""
    name: str
    max_variable_byte_sizes: Tuple[int, ...]
    n_instances: int
    len_lo: int
    len_hi: int
    description: str

    @property
    def blocks_per_instance(self) -> int:
        return sum(m // 64 for m in self.max_variable_byte_sizes)


WORKLOADS = {
    # configs[0]: single 64-byte message, max input 128 bytes (benches/digest.rs shape scaled to 2 blocks)
    "cfg1": Workload("cfg1", (128,), 1, 64, 64, "single 64-byte message, max input 128 bytes"),
    # configs[1]: the configuration the metric is quoted on at N=1
    "cfg2": Workload("cfg2", (64,), 1024, 55, 55, "1024 random 55-byte messages, one block each (max 64)"),
    "cfg3": Workload("cfg3", (1088,), 4096, 0, 1024, "4096 dynamic-length messages 0-1024 bytes, max 1088 (17 blocks)"),
    "cfg4": Workload("cfg4", (320,), 1 << 16, 256, 256, "2^16 messages of 256 bytes, max 320 (5 blocks)"),
    "cfg5": Workload("cfg5", (2112,), 1 << 18, 1, 2048, "2^18 messages uniformly 1-2048 bytes, max 2112 (33 blocks)"),
}


def message_lengths(w: Workload, first: int, count: int, seed: int = SEED) -> np.ndarray:
    idx = np.arange(first, first + count, dtype=np.uint64)
    if w.len_lo == w.len_hi:
        return np.full(count, w.len_lo, dtype=np.uint32)
    r = splitmix64(np.uint64(seed ^ LEN_TAG) + idx)
    return (np.uint64(w.len_lo) + r % np.uint64(w.len_hi - w.len_lo + 1)).astype(np.uint32)


def generate(w: Workload, first: int, count: int, seed: int = SEED):
    """Messages [first, first+count) of workload w -> (blob u8, offsets u64, lens u32), packed back to back."""
    lens = message_lengths(w, first, count, seed)
    max_len = int(lens.max()) if count else 0
    words = (max_len + 7) // 8
    idx = np.arange(first, first + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) + idx * np.uint64(GOLDEN)).astype(np.uint64)
        grid = splitmix64(key[:, None] + np.arange(words, dtype=np.uint64)[None, :]) if words else np.zeros((count, 0), np.uint64)
    rows = grid.view(np.uint8).reshape(count, words * 8) if words else np.zeros((count, 0), np.uint8)
    offs = np.zeros(count, dtype=np.uint64)
    if count > 1:
        offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    if w.len_lo == w.len_hi:
        blob = np.ascontiguousarray(rows[:, :max_len]).reshape(-1)
    else:
        mask = np.arange(words * 8, dtype=np.uint32)[None, :] < lens[:, None]
        blob = rows[mask]
    return np.ascontiguousarray(blob, dtype=np.uint8), offs, lens


def message(w: Workload, m: int, seed: int = SEED) -> bytes:
    blob, _, lens = generate(w, m, 1, seed)
    return bytes(blob[: int(lens[0])])
