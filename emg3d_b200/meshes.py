"""Tensor mesh as consumed by the multigrid path.

Attribute names follow the reference's ``BaseMesh`` (emg3d/meshes.py:42-130) so
that reference meshes and these are interchangeable inside the solver: ``h``,
``origin``, ``shape_cells``, ``shape_nodes``, ``nodes_{x,y,z}``,
``cell_centers_{x,y,z}``, ``shape_edges_{x,y,z}``, ``n_edges*``, ``n_cells``,
``cell_volumes``.  Gridding helpers of the reference (construct_mesh, ...) are
pre-processing and out of scope.
"""
import numpy as np

__all__ = ['BaseMesh', 'TensorMesh']


class BaseMesh:
    """Minimal 3-D tensor-product mesh."""

    def __init__(self, h, origin, **kwargs):
        self.origin = np.array(origin, dtype=float)
        self.h = [np.array(w, dtype=float) for w in h]
        nc = tuple(int(w.size) for w in self.h)
        nn = tuple(n + 1 for n in nc)
        self.shape_cells, self.shape_nodes = nc, nn
        self.n_cells = int(np.prod(nc))
        for ax, name in enumerate('xyz'):
            nodes = np.r_[0., self.h[ax].cumsum()] + self.origin[ax]
            setattr(self, 'nodes_' + name, nodes)
            setattr(self, 'cell_centers_' + name, (nodes[1:] + nodes[:-1]) / 2)
            edges = tuple(nc[a] if a == ax else nn[a] for a in range(3))
            faces = tuple(nn[a] if a == ax else nc[a] for a in range(3))
            setattr(self, 'shape_edges_' + name, edges)
            setattr(self, 'shape_faces_' + name, faces)
            setattr(self, 'n_edges_' + name, int(np.prod(edges)))
            setattr(self, 'n_faces_' + name, int(np.prod(faces)))
        self.n_edges = self.n_edges_x + self.n_edges_y + self.n_edges_z
        self.n_faces = self.n_faces_x + self.n_faces_y + self.n_faces_z

    def __repr__(self):
        return (f"TensorMesh: {self.shape_cells[0]} x {self.shape_cells[1]} x "
                f"{self.shape_cells[2]} ({self.n_cells:,})")

    def __eq__(self, other):
        return (self.shape_cells == getattr(other, 'shape_cells', None) and
                np.allclose(self.origin, other.origin, atol=0) and
                all(np.allclose(a, b, atol=0) for a, b in zip(self.h, other.h)))

    @property
    def cell_volumes(self):
        """Cell volumes as 1-D array, x fastest."""
        if getattr(self, '_cell_volumes', None) is None:
            self._cell_volumes = (
                self.h[0][None, None, :] * self.h[1][None, :, None] *
                self.h[2][:, None, None]).ravel()
        return self._cell_volumes

    def copy(self):
        return type(self)(self.h, self.origin)


class TensorMesh(BaseMesh):
    """Name used by callers of the reference (emg3d/meshes.py:134)."""
