"""Minimal mesh container used when ``trimesh`` is not installed.

The reference wraps marching-cubes output in ``trimesh.Trimesh`` (``utils/reconstruction.py:37``,
``models/reconstruction.py:86``).  ``make_mesh`` returns a real ``trimesh.Trimesh`` when trimesh is
importable, otherwise a :class:`SimpleMesh` exposing the attributes the reference's callers use
(``vertices``, ``faces``, ``export``).  ``SimpleMesh`` does NOT replicate trimesh's ``process=True``
vertex merging.
"""
from __future__ import annotations

import numpy as np


class SimpleMesh:
    def __init__(self, vertices, faces, process=True):
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.asarray(faces).astype(np.int64).reshape(-1, 3)

    def export(self, path: str):
        """Binary little-endian PLY (or ASCII OBJ for ``.obj``)."""
        if path.endswith('.obj'):
            with open(path, 'w') as f:
                for v in self.vertices:
                    f.write('v %.9g %.9g %.9g\n' % tuple(v))
                for t in self.faces + 1:
                    f.write('f %d %d %d\n' % tuple(t))
            return
        header = ('ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n'
                  'property float z\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n'
                  % (len(self.vertices), len(self.faces)))
        faces = np.empty(len(self.faces), dtype=[('n', 'u1'), ('idx', '<i4', (3,))])
        faces['n'] = 3
        faces['idx'] = self.faces
        with open(path, 'wb') as f:
            f.write(header.encode('ascii'))
            f.write(self.vertices.astype('<f4').tobytes())
            f.write(faces.tobytes())

    def __repr__(self):
        return 'SimpleMesh(vertices=%d, faces=%d)' % (len(self.vertices), len(self.faces))


def make_mesh(vertices, faces, process=True):
    try:
        import trimesh  # noqa: WPS433
        if hasattr(trimesh, 'Trimesh'):
            return trimesh.Trimesh(vertices, faces, process=process)
    except ImportError:
        pass
    return SimpleMesh(vertices, faces, process=process)
