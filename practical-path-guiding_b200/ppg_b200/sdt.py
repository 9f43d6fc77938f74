"""Reader of the reference's .sdt SD-tree dumps (written by dumpSDTree, guided_path.cpp:1191-1208 with DTreeWrapper::dump :699-711 and STree::dump :945-951;
by this library: ppg_dump_sdtree / dumpSDTree=true), following the reader of the reference's visualizer (visualizer/src/main.cpp:142-173).

Layout (little endian): 16 x f32 camera-to-world matrix (row major), then for every S-tree leaf with sampling weight > 0, in depth-first order (child 0 first):
3 x f32 min corner, 3 x f32 size, f32 mean radiance, u64 statistical weight, u64 node count, then per quadtree node 4 x (f32 sum, u16 child index; 0 = leaf cell).
Child j of a node covers x in [0.5 (j & 1), ...], y in [0.5 (j >> 1), ...] of the cylindrical map (QuadTreeNode::childIndex, :205-217)."""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

_NODE = np.dtype([("sum", "<f4"), ("child", "<u2")])


@dataclass
class DTreeDump:
    pos: np.ndarray        # (3,) min corner of the leaf's voxel
    size: np.ndarray       # (3,)
    mean: float            # mean radiance of the sampling distribution
    weight: int            # statistical weight (number of records, truncated)
    sums: np.ndarray       # (nodes, 4) float32
    children: np.ndarray   # (nodes, 4) uint16

    def depth(self) -> int:
        """Deepest level of the quadtree (a lone root = 1), visualizer main.cpp: computeDepth."""
        d = np.ones(len(self.sums), np.int64)
        for i in range(len(self.sums) - 1, -1, -1):          # children always follow their parent in the node array
            ch = self.children[i][self.children[i] > 0]
            if len(ch):
                d[i] = 1 + d[ch].max()
        return int(d[0]) if len(d) else 0

    def pdf(self, x: float, y: float) -> float:
        """Density of the cylindrical point (x, y) in [0,1]^2 w.r.t. solid angle (DTree::pdf, :415-421 with :232-245)."""
        if not (self.mean > 0):
            return 1.0 / (4 * np.pi)
        n, res = 0, 1.0
        while True:
            j = (1 if x >= 0.5 else 0) | (2 if y >= 0.5 else 0)
            x = x * 2 - (j & 1); y = y * 2 - (j >> 1)
            tot = float(self.sums[n].sum())
            if not (self.sums[n, j] > 0):
                return 0.0
            res *= 4.0 * float(self.sums[n, j]) / tot
            n = int(self.children[n, j])
            if n == 0:
                return res / (4 * np.pi)


def read(path) -> tuple[np.ndarray, list[DTreeDump]]:
    """(camera-to-world 4x4, the leaves' D-trees)."""
    with open(path, "rb") as f:
        b = f.read()
    if len(b) < 64:
        raise ValueError(f"{path}: too short for an .sdt file")
    cam = np.frombuffer(b, "<f4", 16).reshape(4, 4).copy()
    p, leaves = 64, []
    while p < len(b):
        if p + 44 > len(b):
            raise ValueError(f"{path}: truncated leaf header at byte {p}")
        pos = np.frombuffer(b, "<f4", 3, p).copy(); size = np.frombuffer(b, "<f4", 3, p + 12).copy()
        mean, = struct.unpack_from("<f", b, p + 24); weight, nodes = struct.unpack_from("<QQ", b, p + 28); p += 44
        if p + nodes * 24 > len(b):
            raise ValueError(f"{path}: truncated node array at byte {p}")
        arr = np.frombuffer(b, _NODE, nodes * 4, p).reshape(nodes, 4); p += nodes * 24
        leaves.append(DTreeDump(pos, size, float(mean), int(weight), arr["sum"].copy(), arr["child"].copy()))
    return cam, leaves
