"""The slice of trimesh the reference touches on the tested path: load_mesh(.obj) (dataset/utils.py:19-26) and
Trimesh(vertices, faces).face_normals / .sample(n, return_index=True) (utils/eval_metric.py:38-53)."""
import numpy as np


class Trimesh:
    def __init__(self, vertices=None, faces=None, vertex_colors=None, process=False, **_):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.faces = np.asarray(faces, dtype=np.int64)
        self.vertex_colors = vertex_colors

    @property
    def edges(self):
        return self.faces[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2)

    def _cross(self):
        t = self.vertices[self.faces]
        return np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])

    @property
    def face_normals(self):
        c = self._cross()
        return c / np.maximum(np.linalg.norm(c, axis=1, keepdims=True), 1e-30)

    def sample(self, count, return_index=False):
        area = 0.5 * np.linalg.norm(self._cross(), axis=1)
        idx = np.random.choice(len(area), size=count, p=area / area.sum())
        w = np.random.dirichlet((1,) * 3, count)
        pts = (w[:, :, None] * self.vertices[self.faces[idx]]).sum(axis=1)
        return (pts, idx) if return_index else pts

    def export(self, path, *a, **k):
        with open(path, "w") as f:
            for v in self.vertices:
                f.write("v %.6f %.6f %.6f\n" % tuple(v))
            for t in self.faces + 1:
                f.write("f %d %d %d\n" % tuple(t))


def load_mesh(path, process=False, **_):
    verts, faces = [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                verts.append([float(x) for x in p[1:4]])
            elif p[0] == "f":
                faces.append([int(x.split("/")[0]) - 1 for x in p[1:4]])
    return Trimesh(np.array(verts), np.array(faces))


load = load_mesh
