"""Host-side layout logic of a multi-GPU run (pure Python, no CUDA): which tasks and markers a rank owns and how
window positions of different ranks interleave. Mirrors mpi_define_blocks_of_markers / mpi_assign_blocks_to_tasks
(reference src/BayesRRm.cpp:396-413, 781-827) and the merge rule of the marker kernel."""
from __future__ import annotations

import numpy as np


def define_blocks(m_total: int, n_tasks: int):
    """Contiguous marker blocks; the first m_total % n_tasks tasks get one extra marker."""
    base, modu = divmod(m_total, n_tasks)
    lens = np.array([base + (1 if i < modu else 0) for i in range(n_tasks)], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    return starts, lens


def rank_layout(m_total: int, tasks_total: int, world: int, rank: int):
    """Tasks are dealt to the ranks in contiguous groups of tasks_total // world."""
    assert tasks_total % world == 0, "tasks must divide evenly over the GPUs"
    tl = tasks_total // world
    starts, lens = define_blocks(m_total, tasks_total)
    t0 = rank * tl
    return dict(task_first=t0, tasks_local=tl, m_start=int(starts[t0]), m_local=int(lens[t0:t0 + tl].sum()), lmax=int(lens.max()))


def global_position(p_local: int, tasks_local: int, tasks_total: int, task_first: int) -> int:
    """Window position of a local marker in the order the reference sums its ranks: step * T_total + global task."""
    return (p_local // tasks_local) * tasks_total + task_first + (p_local % tasks_local)


def merge_changed(lists):
    """lists[r] = [(global_position, payload), ...] of rank r -> one list in global window order (what every GPU applies)."""
    out = [e for l in lists for e in l]
    out.sort(key=lambda e: e[0])
    return out
