"""CPU estimate of BVH leaf quality for the two leaf-membership rules of csrc/nodegrid.cu (development aid, numpy only):
runs of eight nodes in Morton order vs a k-d median split along each segment's widest axis.  Prints the mean leaf-box diagonal and the
mean number of leaf boxes a query ball touches (ball radius = the query's true 8th-nearest distance), for near-surface and far queries.
   python tools/bvh_quality.py [--nodes 2036]"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def morton_order(v):
    lo, hi = v.min(0), v.max(0)
    q = np.clip(((v - lo) / (hi - lo).max() * 1023).astype(np.int64), 0, 1023)
    key = np.zeros(len(v), np.int64)
    for b in range(10):
        for a in range(3):
            key |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return np.argsort(key, kind="stable")


def kd_order(v):
    L = 1
    while L * 8 < len(v):
        L *= 2
    n = L * 8
    idx = np.concatenate([np.arange(len(v)), np.full(n - len(v), -1)])

    def rec(seg):
        if len(seg) <= 8:
            return seg
        real = seg[seg >= 0]
        axis = int(np.argmax(v[real].max(0) - v[real].min(0))) if len(real) else 0
        keys = np.where(seg >= 0, v[np.maximum(seg, 0), axis], np.inf)
        o = np.lexsort((np.where(seg >= 0, seg, 2 ** 31 - 1), keys))
        seg = seg[o]
        h = len(seg) // 2
        return np.concatenate([rec(seg[:h]), rec(seg[h:])])
    return rec(idx)


def leaf_boxes(v, order):
    boxes = []
    for l in range(0, len(order), 8):
        m = order[l:l + 8]
        m = m[m >= 0]
        if len(m):
            boxes.append((v[m].min(0), v[m].max(0)))
    return boxes


def touched(boxes, q, r2):
    n = 0
    for lo, hi in boxes:
        d = np.maximum(np.maximum(lo - q, q - hi), 0)
        n += (d * d).sum() <= r2
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=2036)
    a = ap.parse_args()
    from dynamicfusion_b200 import synth
    from oracle import orc
    dim = 160
    depth = synth.umbrella_depth(0)
    vol = np.zeros(dim ** 3, np.uint32)
    vs = (1.0 / dim,) * 3
    pose = synth.volume_pose(1.0)
    ident = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    orc.integrate(vol, (dim,) * 3, vs, 0.04, 64, orc.compute_dists(depth, synth.DEFAULT_K), (pose[0], pose[1]), synth.DEFAULT_K)
    cloud = orc.extract_cloud(vol, (dim,) * 3, vs, 0.04, 64, pose, 4_000_000)[:, :3]
    step = max(1, len(cloud) // a.nodes)
    v = cloud[::step][: a.nodes].astype(np.float64)
    print(f"{len(v)} nodes from a {len(cloud)}-point cloud")
    rng = np.random.default_rng(0)
    near = v[rng.integers(0, len(v), 200)] + rng.normal(scale=0.01, size=(200, 3))
    far = near + np.array([0.0, 0.0, -0.3])
    for name, order in (("morton runs", morton_order(v)), ("k-d split  ", kd_order(v))):
        boxes = leaf_boxes(v, np.asarray(order))
        diag = np.mean([np.linalg.norm(hi - lo) for lo, hi in boxes])
        out = []
        for qs in (near, far):
            cnt = []
            for q in qs:
                d2 = np.sort(((v - q) ** 2).sum(1))[7]
                cnt.append(touched(boxes, q, d2))
            out.append(np.mean(cnt))
        print(f"{name}: {len(boxes)} leaves, mean box diagonal {diag * 1000:.1f} mm, leaves touched per query: near {out[0]:.1f}, far (0.3 m off the surface) {out[1]:.1f}")


if __name__ == "__main__":
    main()
