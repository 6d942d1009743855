"""CPU restatement of RayJoin's polygon overlay (test infrastructure only).

Follows, on top of the C oracle's LSI / PIP:
  IntersectEdge(0)                       src/run_overlay.cu:206
  LocateVerticesInOtherMap(im)           src/app/map_overlay_lbvh.h:73-107
  ComputeOutputPolygons                  src/app/map_overlay_lbvh.h:109-265
  WriteOutputChain                       src/app/output_chain.h:41-205
Pure Python loops: meant for maps of a few thousand edges.
"""
import numpy as np

from . import oracle as O

DONTKNOW = -1


class OverlayOracle:
    def __init__(self, graphs, bbox=None, grid_size=None):
        """grid_size: None = the LBVH / RT pair set (pure predicate); a number = the pair set of
        the reference's GRID backend at that -grid_size (src/app/lsi_grid.h:62-67: a pair is
        kept only in the cell of its intersection point)."""
        self.grid_size = grid_size
        from rayjoin_b200 import synth  # only the bbox helper, no compute
        self.g = graphs
        self.bbox = bbox or synth.union_bbox(*graphs)
        self.sc = O.scaling_init(*self.bbox)
        self.pts = [O.scale_points(self.sc, g.xy) for g in graphs]
        e = [O.build_edges(g.row_index) for g in graphs]
        self.p1 = [x[0] for x in e]
        self.chain = [x[1] for x in e]

    def _pip_faces(self, q, pts):
        b = 1 - q
        eids = O.pip_grid(self.pts[b], self.p1[b], self.sc, pts, q)
        faces = O.face_ids(self.pts[b], self.p1[b], self.chain[b], self.g[b].left, self.g[b].right, eids)
        return eids, faces

    def run(self):
        # LSI with map 0 as the query side: pairs (eid0, eid1, x, y)
        if self.grid_size:
            e0, e1, x, y = O.lsi_refgrid(self.pts[0], self.p1[0], self.pts[1], self.p1[1], self.sc,
                                         self.grid_size, sort_map=0)
        else:
            e0, e1, x, y = O.lsi_grid(self.pts[0], self.p1[0], self.pts[1], self.p1[1], self.sc)
        self.xs = np.column_stack([e0.astype(np.int64), e1.astype(np.int64), x, y])
        self.closest = [None, None]
        self.pip = [None, None]
        for im in range(2):
            self.closest[im], self.pip[im] = self._pip_faces(im, self.pts[im])
        self.sorted = [None, None]   # rows: eid0, eid1, x, y, mid_point_polygon_id
        for im in range(2):
            rows = [list(map(int, r)) + [DONTKNOW] for r in self.xs]
            groups = {}
            for r in rows:
                groups.setdefault(r[im], []).append(r)
            out, mids, owners = [], [], []
            for eid in sorted(groups):
                grp = groups[eid]
                p = self.pts[im][self.p1[im][eid]]
                px, py = int(p[0]), int(p[1])
                # squared distance to p1 in exact integers (map_overlay_lbvh.h:204-214);
                # ties by the other map's eid (the reference leaves them unordered)
                grp.sort(key=lambda r: ((r[2] - px) ** 2 + (r[3] - py) ** 2, r[1 - im]))
                for a, b in zip(grp[:-1], grp[1:]):
                    # x1 + (x2 - x1) / 2 in rationals, truncated toward zero (:216-228)
                    mids.append((_tdiv2(a[2] + b[2]), _tdiv2(a[3] + b[3])))
                    owners.append(a)
                out.extend(grp)
            if mids:
                _, faces = self._pip_faces(im, np.asarray(mids, np.int64))
                for r, f in zip(owners, faces):
                    r[4] = int(f)
            self.sorted[im] = out
        return self

    def write(self, path):
        sc = self.sc
        chains = []

        def unscale(r):
            tx = float(np.float64(r[2]) * np.float64(sc.rrx))
            ty = float(np.float64(r[3]) * np.float64(sc.rry))
            return (float(np.float64(tx) + np.float64(sc.ddeltax)),
                    float(np.float64(ty) + np.float64(sc.ddeltay)))

        for im in range(2):
            g = self.g[im]
            groups = {}
            for r in self.sorted[im]:
                groups.setdefault(r[im], []).append(r)
            pip = self.pip[im]
            for ic in range(g.n_chains):
                cur = {"pts": [], "left": int(g.left[ic]), "right": int(g.right[ic]), "other": 0}

                def flush():
                    if cur["pts"]:
                        if cur["left"] * cur["other"] != 0 or cur["right"] * cur["other"] != 0:
                            pts = [cur["pts"][0]]
                            for p in cur["pts"][1:]:
                                if p != pts[-1]:
                                    pts.append(p)
                            chains.append({"pts": pts, "left": cur["left"], "right": cur["right"],
                                           "other": cur["other"]})
                        cur["pts"] = []
                b, e = int(g.row_index[ic]), int(g.row_index[ic + 1])
                for pid in range(b, e):
                    cur["other"] = int(pip[pid])
                    cur["pts"].append((float(g.xy[pid, 0]), float(g.xy[pid, 1])))
                    if pid != e - 1:
                        grp = groups.get(pid - ic)
                        if grp:
                            cur["pts"].append(unscale(grp[0]))
                            for a, nb in zip(grp[:-1], grp[1:]):
                                flush()
                                cur["other"] = a[4]
                                cur["pts"].append(unscale(a))
                                cur["pts"].append(unscale(nb))
                            flush()
                            cur["pts"].append(unscale(grp[-1]))
                flush()
        face_ids, point_ids = {}, {}

        def create_polygon(a, b):
            if a == 0 or b == 0:
                return 0
            if (a, b) not in face_ids:
                face_ids[(a, b)] = len(face_ids) + 1
            return face_ids[(a, b)]
        lines = []
        for ch in chains:
            o = ch["other"]
            ch["left"] = create_polygon(ch["left"], o) if ch["left"] < o else create_polygon(o, ch["left"])
            ch["right"] = create_polygon(ch["right"], o) if ch["right"] < o else create_polygon(o, ch["right"])
            for p in ch["pts"]:
                if p not in point_ids:
                    point_ids[p] = len(point_ids)
        for i, ch in enumerate(chains):
            lines.append("%d %d %d %d %d %d" % (i + 1, len(ch["pts"]), point_ids[ch["pts"][0]],
                                                point_ids[ch["pts"][-1]], ch["left"], ch["right"]))
            for p in ch["pts"]:
                lines.append("%.6f %.6f" % p)
        with open(path, "w") as f:
            f.write("\n".join(lines) + ("\n" if lines else ""))
        return len(chains)


def _tdiv2(v):
    """C-style division by two (truncation toward zero)."""
    return -((-v) // 2) if v < 0 else v // 2
