"""Structured voxel grid numbering (oracle side; test infrastructure only).

Restates the index contract of ``pymoto/common/domain.py`` (VoxelDomain):
  element number  (k*nely + j)*nelx + i                       domain.py:200-211
  node number     (k*(nely+1) + j)*(nelx+1) + i               domain.py:213-224
  dof number      node*ndof + d                               domain.py:226-251
  local node a    bit0 -> +x, bit1 -> +y, bit2 -> +z          domain.py:132-142, 338-351
"""
import numpy as np


class Grid:
    def __init__(self, nelx, nely, nelz=0, unitx=1.0, unity=1.0, unitz=1.0):
        self.nelx, self.nely, self.nelz = int(nelx), int(nely), int(nelz or 0)
        self.unitx, self.unity, self.unitz = float(unitx), float(unity), float(unitz)
        self.dim = 2 if self.nelz == 0 else 3
        self.nel = self.nelx * self.nely * max(self.nelz, 1)
        self.nnodes = (self.nelx + 1) * (self.nely + 1) * (self.nelz + 1)
        self.elemnodes = 2 ** self.dim
        # sign pattern of local node a along (x, y, z): -1 / +1, x fastest (domain.py:132-142)
        self.node_signs = np.array(
            [[(1 if (a >> b) & 1 else -1) if b < self.dim else 0 for b in range(3)] for a in range(self.elemnodes)]
        )

    @property
    def element_size(self):
        return np.array([self.unitx, self.unity, self.unitz])

    @property
    def size(self):
        return np.array([self.nelx, self.nely, self.nelz])[: self.dim]

    def coarsen(self):
        """The sub-domain of one multigrid level (solvers/iterative.py:162-164)."""
        return Grid(self.nelx // 2, self.nely // 2, self.nelz // 2, self.unitx * 2, self.unity * 2, self.unitz * 2)

    def elem_number(self, i, j, k=0):
        return (k * self.nely + j) * self.nelx + i

    def node_number(self, i, j, k=0):
        return (k * (self.nely + 1) + j) * (self.nelx + 1) + i

    def node_indices(self, n):
        n = np.asarray(n)
        return n % (self.nelx + 1), (n // (self.nelx + 1)) % (self.nely + 1), n // ((self.nelx + 1) * (self.nely + 1))

    def elem_indices(self, e):
        e = np.asarray(e)
        return e % self.nelx, (e // self.nelx) % self.nely, e // (self.nelx * self.nely)

    def nodes3d(self):
        """Array [i, j, k] -> node number, like ``VoxelDomain.nodes`` (domain.py:160-164)."""
        i, j, k = np.meshgrid(np.arange(self.nelx + 1), np.arange(self.nely + 1), np.arange(self.nelz + 1), indexing="ij")
        return self.node_number(i, j, k)

    def conn(self):
        """(nel, elemnodes) node numbers of every element in element-number order (domain.py:144-152, 338-351)."""
        e = np.arange(self.nel)
        i, j, k = self.elem_indices(e)
        cols = []
        for a in range(self.elemnodes):
            cols.append(self.node_number(i + (a & 1), j + ((a >> 1) & 1), k + ((a >> 2) & 1)))
        return np.stack(cols, axis=-1)

    def dofconn(self, ndof):
        """(nel, elemnodes*ndof) dof numbers, node-major then dof (domain.py:353-362)."""
        c = self.conn()
        return (c[:, :, None] * ndof + np.arange(ndof)[None, None, :]).reshape(self.nel, -1)

    # ---- shape functions (domain.py:364-428) ----
    def shape_fun_der(self, pos):
        v = np.prod(self.element_size[: self.dim])
        dN = np.ones((self.dim, self.elemnodes)) / v
        sg = self.node_signs
        for i in range(self.dim):
            for j in range(self.dim):
                if i != j:
                    dN[i, :] *= self.element_size[j] / 2 + sg[:, j] * pos[j]
            dN[i, :] *= sg[:, i]
        return dN
