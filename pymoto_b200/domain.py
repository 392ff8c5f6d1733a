"""Structured voxel grid: the index contract of the hot path.

Mirrors the numbering API of the reference's ``VoxelDomain`` (pymoto/common/domain.py:57-428; alias
``DomainDefinition``, pymoto/__init__.py:102-117): element number ``(k*nely + j)*nelx + i`` (:211), node number
``(k*(nely+1) + j)*(nelx+1) + i`` (:224), dof number ``node*ndof + d`` (:247), local node order with x the
fastest bit (:132-142).  Any object with ``nelx, nely, nelz, unitx, unity, unitz`` (e.g. a real
``pymoto.VoxelDomain``) is accepted by the modules of this package; the big index arrays (``conn``, ``nodes``,
``elements``) are built lazily because the CUDA kernels never need them (all indexing is closed-form).
"""
import numpy as np


class VoxelDomain:
    def __init__(self, nelx: int, nely: int, nelz: int = 0, unitx: float = 1.0, unity: float = 1.0, unitz: float = 1.0):
        self.nelx, self.nely, self.nelz = int(nelx), int(nely or 0), int(nelz or 0)
        self.unitx, self.unity, self.unitz = float(unitx), float(unity), float(unitz)
        if self.nelx < 1 or self.nely < 1:
            raise ValueError("pymoto_b200.VoxelDomain supports 2-D and 3-D grids (nelx, nely >= 1)")
        self.dim = 2 if self.nelz == 0 else 3
        assert np.prod(self.element_size[: self.dim]) > 0.0, "Element volume needs to be positive"
        self.origin = np.zeros(3)
        self.nel = self.nelx * self.nely * max(self.nelz, 1)
        self.nnodes = (self.nelx + 1) * (self.nely + 1) * (self.nelz + 1)
        self.elemnodes = 2 ** self.dim
        self.node_numbering = [[(1 if (a >> b) & 1 else -1) if b < self.dim else 0 for b in range(3)]
                               for a in range(self.elemnodes)]
        self._conn = self._nodes = self._elements = None

    # ---- sizes
    @property
    def element_size(self):
        return np.array([self.unitx, self.unity, self.unitz])

    @property
    def domain_size(self):
        return np.array([self.nelx * self.unitx, self.nely * self.unity, self.nelz * self.unitz])[: self.dim]

    @property
    def size(self):
        return np.array([self.nelx, self.nely, self.nelz])[: self.dim]

    # ---- numbering
    def get_elemnumber(self, eli, elj, elk=0):
        return (elk * self.nely + elj) * self.nelx + eli

    def get_nodenumber(self, nodi, nodj, nodk=0):
        return (nodk * (self.nely + 1) + nodj) * (self.nelx + 1) + nodi

    def get_dofnumber(self, nod_idx, dof_idx=None, ndof=None):
        ndof = self.dim if ndof is None else ndof
        nod = nod_idx if isinstance(nod_idx, int) else np.asarray(nod_idx)
        dof = np.arange(ndof) if dof_idx is None else (dof_idx if isinstance(dof_idx, int) else np.asarray(dof_idx))
        if np.ndim(dof) == 0 or np.ndim(nod) == 0:
            return nod * ndof + dof
        nod = nod.reshape(nod.shape + (1,) * np.ndim(dof))
        return nod * ndof + dof

    def get_node_indices(self, nod_idx=None):
        n = np.arange(self.nnodes) if nod_idx is None else np.asarray(nod_idx)
        i = n % (self.nelx + 1)
        j = (n // (self.nelx + 1)) % (self.nely + 1)
        if self.dim == 2:
            return np.stack([i, j], axis=0)
        return np.stack([i, j, n // ((self.nelx + 1) * (self.nely + 1))], axis=0)

    def get_element_indices(self, el_idx=None):
        e = np.arange(self.nel) if el_idx is None else np.asarray(el_idx)
        i = e % self.nelx
        j = (e // self.nelx) % self.nely
        if self.dim == 2:
            return np.stack([i, j], axis=0)
        return np.stack([i, j, e // (self.nelx * self.nely)], axis=0)

    def get_node_position(self, nod_idx=None):
        ijk = self.get_node_indices(nod_idx)
        return (self.origin[: self.dim] + self.element_size[: self.dim] * ijk.T).T

    def get_elemconnectivity(self, i, j, k=0):
        return np.stack([self.get_nodenumber(i + (a & 1), j + ((a >> 1) & 1), k + ((a >> 2) & 1))
                         for a in range(self.elemnodes)], axis=-1)

    def get_dofconnectivity(self, ndof: int):
        return self.get_dofnumber(self.conn, ndof=ndof).reshape(self.nel, -1)

    # ---- lazily materialised index arrays
    @property
    def conn(self):
        if self._conn is None:
            ijk = self.get_element_indices()
            k = ijk[2] if self.dim == 3 else 0
            self._conn = self.get_elemconnectivity(ijk[0], ijk[1], k)
        return self._conn

    @property
    def nodes(self):
        if self._nodes is None:
            i, j, k = np.meshgrid(np.arange(self.nelx + 1), np.arange(self.nely + 1), np.arange(self.nelz + 1), indexing="ij")
            self._nodes = self.get_nodenumber(i, j, k)
        return self._nodes

    @property
    def elements(self):
        if self._elements is None:
            i, j, k = np.meshgrid(np.arange(self.nelx), np.arange(self.nely), np.arange(max(self.nelz, 1)), indexing="ij")
            self._elements = self.get_elemnumber(i, j, k)
        return self._elements

    # ---- trilinear shape functions on [-h/2, h/2]^dim (domain.py:364-428)
    def eval_shape_fun(self, pos):
        h = self.element_size
        sg = np.array(self.node_numbering, dtype=float)
        N = np.full(self.elemnodes, 1.0 / np.prod(h[: self.dim]))
        for a in range(self.dim):
            N *= h[a] / 2 + sg[:, a] * pos[a]
        return N

    def eval_shape_fun_der(self, pos):
        h = self.element_size
        sg = np.array(self.node_numbering, dtype=float)
        dN = np.ones((self.dim, self.elemnodes)) / np.prod(h[: self.dim])
        for a in range(self.dim):
            for b in range(self.dim):
                if a != b:
                    dN[a, :] *= h[b] / 2 + sg[:, b] * pos[b]
            dN[a, :] *= sg[:, a]
        return dN


DomainDefinition = VoxelDomain  # deprecated name used by the north star (pymoto/__init__.py:102-117)


def grid_dims(domain):
    """(nelx, nely, nelz) of any domain-like object."""
    return int(domain.nelx), int(domain.nely), int(getattr(domain, "nelz", 0) or 0)
