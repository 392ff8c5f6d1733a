"""DeviceDyad: the matrix sensitivity dg/dK = sum_k u_k (x) v_k kept as device vectors.

Counterpart of the reference's ``DyadicMatrix`` (pymoto/common/dyadcarrier.py:34-525) for the one use the hot
path makes of it: ``LinSolve._sensitivity`` returns ``DyadicMatrix(-lam, u)`` (pymoto/modules/linalg.py:204-209)
and ``AssembleGeneral._sensitivity`` zeroes the Dirichlet rows/columns and contracts it with the element matrix
(pymoto/modules/assembly.py:301-314).  The zeroing and the contraction happen inside ``pmb_assemble_sens``.
"""
import torch


class DeviceDyad:
    def __init__(self, u=None, v=None):
        self.u = [] if u is None else (list(u) if isinstance(u, (list, tuple)) else [u])
        self.v = [] if v is None else (list(v) if isinstance(v, (list, tuple)) else [v])
        if len(self.u) != len(self.v):
            raise TypeError("Number of vectors in u and v must be equal")

    @property
    def size(self):
        return 0 if not self.u else self.u[0].numel() * self.v[0].numel()

    @property
    def shape(self):
        return (-1, -1) if not self.u else (self.u[0].numel(), self.v[0].numel())

    @property
    def n_dyads(self):
        return len(self.u)

    @property
    def real(self):
        return self

    def add_dyad(self, u, v):
        self.u.append(u)
        self.v.append(v)
        return self

    def __iadd__(self, other):
        if not isinstance(other, DeviceDyad):
            return NotImplemented
        self.u += other.u
        self.v += other.v
        return self

    def __deepcopy__(self, memo):
        return DeviceDyad([t.clone() for t in self.u], [t.clone() for t in self.v])

    def todense(self):
        """Dense numpy matrix (tests on small problems only)."""
        out = 0
        for u, v in zip(self.u, self.v):
            out = out + torch.outer(u, v)
        return out.cpu().numpy()
