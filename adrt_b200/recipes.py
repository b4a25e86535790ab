"""Device-resident solvers built on the hot path.

The reference ships these only as documentation recipes
(/root/reference/docs/examples.cginverse.md:40-67 ``ADRTNormalOperator`` /
``iadrt_cg``, docs/examples.tomography.md:63-70 ridge variant): a conjugate
gradient solve of ``A^T A x = A^T b`` with ``A = adrt`` and
``A^T = mean_q(truncate(bdrt(.)))``.  Here every CG iteration stays on the GPU: the
operator is ONE native call (``adrt_b200_normal_operator``: forward passes, back-projection
restricted to the offsets ``truncate`` keeps, ``truncate_mean``; the sinogram passes from
adrt to bdrt as workspace rows and never takes the public layout); the vector updates are three
native passes with their dot products (``adrt_b200_cg_dot`` / ``_cg_update`` / ``_cg_direction``).
"""
from __future__ import annotations

import numpy as np

from . import _adrt_cdefs as cd
from ._wrappers import _normalize_array
from .core import _to_device


def normal_operator(x, /, *, ridge: float = 0.0, dist=None):
    """``mean_q(truncate(bdrt(adrt(x)))) + ridge * x`` for CUDA tensors ``(B?, n, n)``.

    With an initialised ``torch.distributed`` module passed as `dist` (and `x`
    replicated on the ranks) the four quadrants are computed on different GPUs
    and exchanged with one all-gather (``_shard.sharded_normal_operator``); the
    result is bit-identical to the single-GPU one."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        from ._shard import sharded_normal_operator

        out = sharded_normal_operator(x, dist)
    else:
        out = cd.normal_operator(x, 1.0)
    if ridge:
        out = out + ridge * x
    return out


def iadrt_cg(b, /, *, ridge: float = 0.0, rtol: float = 1e-5, atol: float = 0.0, maxiter=None, x0=None,
             return_info: bool = False, dist=None):
    """Inverse ADRT by conjugate gradients on the normal equations.

    ``b``: ADRT-shaped array ``(4, 2n-1, n)`` (NumPy or CUDA tensor, no batch
    dimension, like the reference recipe).  Stops when ``||r|| <= max(rtol*||A^T b||, atol)``
    (SciPy's ``cg`` criterion) or after ``maxiter`` iterations (default ``10 n^2``
    like SciPy); raises ``ValueError`` if it did not converge, as the recipe does.
    `dist`: pass ``torch.distributed`` to shard every operator application over
    the ranks by quadrant (`b` replicated; every rank returns the same image).
    """
    import torch

    if b.ndim > 3:
        raise ValueError("batch dimension not supported for iadrt_cg")
    b = _normalize_array(b)
    as_numpy = isinstance(b, np.ndarray)
    bt = _to_device(b) if as_numpy else b
    n = bt.shape[-1]
    rhs = cd.bdrt_truncate_mean(bt, 1.0).contiguous()
    x = torch.zeros_like(rhs) if x0 is None else (_to_device(x0) if isinstance(x0, np.ndarray) else x0).clone().contiguous()
    r = (rhs - normal_operator(x, ridge=ridge, dist=dist)).contiguous() if x0 is not None else rhs.clone()
    p = r.clone()
    tol = max(rtol * float(torch.linalg.vector_norm(rhs)), atol)
    maxiter = 10 * n * n if maxiter is None else int(maxiter)
    # the vector updates of an iteration are three native passes (adrt_b200_cg_dot / _cg_update / _cg_direction);
    # the scalars r.r, p.Ap and the new r.r stay on the device, only the stopping test reads one back
    from . import _lib

    lib = _lib.load()
    code = _lib.F32 if rhs.dtype == torch.float32 else _lib.F64
    count = rhs.numel()
    with torch.cuda.device(rhs.device):
        state = torch.zeros(4, dtype=torch.float64, device=rhs.device)
        wsb = int(lib.adrt_b200_cg_workspace_bytes())
        ws = torch.empty(wsb, dtype=torch.uint8, device=rhs.device)

        def stream():
            return torch.cuda.current_stream().cuda_stream

        _lib.check(lib.adrt_b200_cg_dot(r.data_ptr(), r.data_ptr(), state.data_ptr(), 0, count, code, ws.data_ptr(), wsb, stream()), "cg_dot")
        it, converged = 0, float(state[0]) ** 0.5 <= tol
        while not converged and it < maxiter:
            ap = normal_operator(p, ridge=ridge, dist=dist).contiguous()
            _lib.check(lib.adrt_b200_cg_dot(p.data_ptr(), ap.data_ptr(), state.data_ptr(), 1, count, code, ws.data_ptr(), wsb, stream()), "cg_dot")
            _lib.check(lib.adrt_b200_cg_update(x.data_ptr(), r.data_ptr(), p.data_ptr(), ap.data_ptr(), state.data_ptr(), count, code,
                                               ws.data_ptr(), wsb, stream()), "cg_update")
            it += 1
            if float(state[2]) ** 0.5 <= tol:
                converged = True
                break
            _lib.check(lib.adrt_b200_cg_direction(p.data_ptr(), r.data_ptr(), state.data_ptr(), count, code, stream()), "cg_direction")
    if not converged:
        raise ValueError(f"convergence failed (cg status {it})")
    out = x.cpu().numpy() if as_numpy else x
    return (out, it) if return_info else out
