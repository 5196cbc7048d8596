#!/usr/bin/env python
"""Offline tree-quality lab (CPU only): how many node / triangle visits does a ray cost in the Morton-order LBVH, and how many after the
opt-in binned-SAH rebuild of its inner nodes (csrc/lbvh.cu: bvh2_sah_rebuild_host, B200PT_BVH_SAH=1)?

The LBVH here is a numpy MIRROR of the GPU builder's recipe (csrc/lbvh.cu: split clipping of fat slivers, 60-bit Morton keys + 2 size-class
bits, Karras topology = split at the highest differing key bit, leaves of <= 4 consecutive slots), not the GPU build itself -- node counts
agree with the GPU build to within the rounding of the keys; treat the visit counts as a model.  The traversal model follows
bvh_traverse.cuh: pop a node, slab-test both child boxes against [0, t_closest], descend into the nearer child first, test a leaf's
triangles when it is reached.  Rays: camera rays of the scene's own camera plus one cosine-distributed bounce ray from every hit.

usage: python profiles/bvh_lab.py tests/golden/breakfast_room.npz [--rays 20000] [--out profiles/r01_bvh_lab_breakfast.json]
"""
import argparse
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vpt_b200 import binding as B                                        # noqa: E402  (host-only entry points: no GPU needed)

LEAF_MAX, SPLIT_MAX = 4, 16


def load_fixture(path):
    """tests/golden/*.npz scene fixture (flat arrays, see tests/golden/make_golden.py) or any file the product's loader reads."""
    if path.endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        meshes = [(np.asarray(z[f"mesh{i}_v"], np.float32)[:, :3], np.asarray(z[f"mesh{i}_i"])) for i in range(int(z["n_meshes"]))]
        inst = [(z["inst_xf"][i], int(z["inst_mesh"][i])) for i in range(len(z["inst_mesh"]))]
        return dict(meshes=meshes, instances=inst, camera_view=np.asarray(z["camera_view"], np.float32), aspect=float(z["aspect"]))
    sc = B.load_gltf(path)                                               # the product's own C++ loader
    meshes = [(np.ascontiguousarray(v["pos"], np.float32), np.asarray(i)) for v, i in sc["meshes"]]
    return dict(meshes=meshes, instances=[(x, m) for x, m, _ in sc["instances"]], camera_view=sc["camera_view"], aspect=float(sc["aspect"]))


def world_triangles(sc):
    out = []
    for xf, mi in sc["instances"]:
        P0, idx = sc["meshes"][mi]
        M = np.asarray(xf, np.float32).reshape(4, 4).T                   # column-major 4x4
        P = P0 @ M[:3, :3].T + M[:3, 3]
        out.append(P[idx.reshape(-1, 3)])
    return np.concatenate(out).astype(np.float32)                        # [T, 3, 3]


def split_refs(tri):
    """(ref boxes [R, 2, 3], ref -> triangle) following k_split_count / k_make_refs."""
    lo, hi = tri.min(1), tri.max(1)
    e = np.sort(hi - lo, axis=1)[:, ::-1]
    area2 = np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    scene = (hi.max(0) - lo.min(0)).max()
    ratio = e[:, 0] * e[:, 1] / np.maximum(area2, 1e-30)
    k = np.ones(len(tri), np.int64)
    m = (ratio > 8.0) & (e[:, 0] > scene / 1024.0)
    k[m] = np.clip(np.ceil(np.sqrt(ratio[m])), 2, SPLIT_MAX).astype(np.int64)
    boxes, owner = [np.stack([lo[k == 1], hi[k == 1]], 1)], [np.nonzero(k == 1)[0]]
    axis = np.argmax(hi - lo, axis=1)
    for j in range(SPLIT_MAX):
        sel = np.nonzero(k > j)[0]; sel = sel[k[sel] > 1]
        if not len(sel): break
        a = axis[sel]; T = tri[sel]; kk = k[sel].astype(np.float32)
        l0 = lo[sel, a]; w = (hi[sel, a] - l0) / kk
        s0 = l0 + w * j; s1 = np.where(j + 1 == k[sel], hi[sel, a], l0 + w * (j + 1))
        ca = np.take_along_axis(T, a[:, None, None].repeat(3, 1), 2)[:, :, 0]          # [n, 3] coordinate along the split axis
        pts, valid = [T], [(ca >= s0[:, None]) & (ca <= s1[:, None])]
        for plane in (s0, s1):
            for u, v in ((0, 1), (1, 2), (2, 0)):
                pu, pv = ca[:, u], ca[:, v]
                cross = (pu < plane) != (pv < plane)
                t = np.where(cross, (plane - pu) / np.where(pv != pu, pv - pu, 1.0), 0.0)
                pts.append((T[:, u] + t[:, None] * (T[:, v] - T[:, u]))[:, None, :]); valid.append(cross[:, None])
        P = np.concatenate(pts, 1); V = np.concatenate(valid, 1)
        blo = np.where(V[:, :, None], P, np.inf).min(1); bhi = np.where(V[:, :, None], P, -np.inf).max(1)
        pad = 1e-5 * np.maximum(np.abs(blo), np.abs(bhi)) + 1e-6 * (hi[sel] - lo[sel])
        L = lo[sel].copy(); H = hi[sel].copy()
        L[np.arange(len(sel)), a] = s0; H[np.arange(len(sel)), a] = s1
        ok = np.isfinite(blo).all(1)
        L2 = np.maximum(L, blo - pad); H2 = np.minimum(H, bhi + pad)
        good = ok[:, None] & (L2 <= H2)
        L = np.where(good, L2, L); H = np.where(good, H2, H)
        boxes.append(np.stack([L, H], 1).astype(np.float32)); owner.append(sel)
    return np.concatenate(boxes), np.concatenate(owner)


def expand21(x):
    x = x & np.uint64(0x1fffff)
    for sh, mask in ((32, 0x1f00000000ffff), (16, 0x1f0000ff0000ff), (8, 0x100f00f00f00f00f), (4, 0x10c30c30c30c30c3), (2, 0x1249249249249249)):
        x = (x | (x << np.uint64(sh))) & np.uint64(mask)
    return x


def morton_keys(rb, size_class=False):                 # LBVH_SIZE_CLASS is 0 in the product build (profiles/r01_variants.txt: rejected)
    lo, hi = rb[:, 0].min(0), rb[:, 1].max(0)
    ext = hi - lo
    c = 0.5 * (rb[:, 0] + rb[:, 1])
    f = np.where(ext > 0, (c - lo) / np.where(ext > 0, ext, 1), 0.0).astype(np.float32)
    q = np.clip(f * np.float32(1048576.0), 0, 1048575).astype(np.uint64)
    rel = np.where(ext > 0, (rb[:, 1] - rb[:, 0]) / np.where(ext > 0, ext, 1), 0.0).max(1)
    cls = np.zeros(len(rb), np.uint64)
    if size_class:
        cls[rel > 0.015625] = 1; cls[rel > 0.0625] = 2; cls[rel > 0.25] = 3
    return (cls << np.uint64(60)) | (expand21(q[:, 0]) << np.uint64(2)) | (expand21(q[:, 1]) << np.uint64(1)) | expand21(q[:, 2])


def build_lbvh(rb, keys):
    """Karras topology over the sorted keys, leaves of <= LEAF_MAX slots; returns BVH2 nodes (root 0)."""
    n = len(keys)
    kl = [int(x) for x in keys]                                          # Python ints: 62-bit arithmetic without overflow games
    nodes = np.zeros(max(n - 1, 1), B.BVH2_DTYPE)
    lo_pref = np.minimum.accumulate  # noqa: F841
    n_out = 0
    # boxes of ranges through a sparse table would be exact; a segment reduce per node is simpler and fast enough (n log n total)
    stack = [(0, n - 1, -1, 0)]
    import bisect
    while stack:
        l, r, parent, side = stack.pop()
        me = n_out; n_out += 1
        if parent >= 0: nodes[parent]["c1" if side else "c0"] = me
        a, b = kl[l], kl[r]
        if a != b:
            d = (a ^ b).bit_length() - 1
            first_with_bit = ((a >> (d + 1)) << (d + 1)) | (1 << d)
            split = bisect.bisect_left(kl, first_with_bit, l, r + 1) - 1   # last slot of the left half
        else:
            d = (l ^ r).bit_length() - 1                                 # equal keys: Karras' index tie-break
            split = (((l >> (d + 1)) << (d + 1)) | (1 << d)) - 1
        for s, (x, y) in enumerate(((l, split), (split + 1, r))):
            nodes[me]["lo1" if s else "lo0"] = rb[x:y + 1, 0].min(0); nodes[me]["hi1" if s else "hi0"] = rb[x:y + 1, 1].max(0)
            if y - x + 1 <= LEAF_MAX: nodes[me]["c1" if s else "c0"] = ~((x << 2) | (y - x))
            else: stack.append((x, y, me, s))
    return nodes[:n_out].copy()


def tree_depth(nodes):
    dep = np.zeros(len(nodes), np.int32); dep[0] = 1
    for i in range(len(nodes)):                                          # both layouts used here store children after their parent
        for c in (int(nodes[i]["c0"]), int(nodes[i]["c1"])):
            if c >= 0: dep[c] = dep[i] + 1
    return int(dep.max())


def traverse(nodes, tris, O, D, tmax=None):
    """Vectorised closest-hit walk.  Returns (t, node visits, triangle tests) per ray."""
    R = len(O); inv = 1.0 / np.where(np.abs(D) > 1e-20, D, 1e-20)
    t = np.full(R, np.inf if tmax is None else tmax, np.float64)
    n_vis = np.zeros(R, np.int64); t_tests = np.zeros(R, np.int64)
    depth = 96
    stack = np.zeros((R, depth), np.int64); sp = np.ones(R, np.int64); stack[:, 0] = 0
    lo0, hi0, lo1, hi1 = (nodes[k].astype(np.float64) for k in ("lo0", "hi0", "lo1", "hi1"))
    c0, c1 = nodes["c0"].astype(np.int64), nodes["c1"].astype(np.int64)
    A, E1, E2 = tris[:, 0].astype(np.float64), (tris[:, 1] - tris[:, 0]).astype(np.float64), (tris[:, 2] - tris[:, 0]).astype(np.float64)

    def slab(lo, hi, o, iv, tt):
        a = (lo - o) * iv; b = (hi - o) * iv
        tn = np.minimum(a, b).max(1); tf = np.maximum(a, b).min(1)
        tn = np.maximum(tn, 0.0)
        return (tn <= np.minimum(tf, tt)), tn

    def leaf(rays, ref):
        r = ~ref; first = r >> 2; cnt = (r & 3) + 1
        for j in range(4):
            m = cnt > j
            if not m.any(): break
            ri = rays[m]; ti = first[m] + j
            t_tests[ri] += 1
            o, d = O[ri], D[ri]
            p = np.cross(d, E2[ti]); det = (E1[ti] * p).sum(1)
            ok = np.abs(det) > 1e-30
            idet = 1.0 / np.where(ok, det, 1.0)
            s = o - A[ti]; u = (s * p).sum(1) * idet
            q = np.cross(s, E1[ti]); v = (d * q).sum(1) * idet
            tt = (E2[ti] * q).sum(1) * idet
            hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (tt > 1e-6) & (tt < t[ri])
            t[ri[hit]] = tt[hit]

    active = np.arange(R)
    while len(active):
        sp[active] -= 1
        n = stack[active, sp[active]]
        n_vis[active] += 1
        o, iv, tt = O[active], inv[active], t[active]
        h0, n0 = slab(lo0[n], hi0[n], o, iv, tt); h1, n1 = slab(lo1[n], hi1[n], o, iv, tt)
        k0, k1 = c0[n], c1[n]
        # leaves are tested at once (nearer first), inner children are pushed far-first
        first_is_0 = n0 <= n1
        for near_pass in (True, False):
            for which, h, k in ((0, h0, k0), (1, h1, k1)):
                is_near = first_is_0 if which == 0 else ~first_is_0
                m = h & (k < 0) & (is_near == near_pass)
                if m.any():
                    if not near_pass:                                    # the nearer leaf may have shortened the ray
                        hh, _ = slab((lo0 if which == 0 else lo1)[n[m]], (hi0 if which == 0 else hi1)[n[m]], O[active[m]], inv[active[m]], t[active[m]])
                        mm = np.nonzero(m)[0][hh]
                        leaf(active[mm], k[mm])
                    else:
                        leaf(active[m], k[m])
        for near_pass in (False, True):                                  # push far first, near last
            for which, h, k in ((0, h0, k0), (1, h1, k1)):
                is_near = first_is_0 if which == 0 else ~first_is_0
                m = h & (k >= 0) & (is_near == near_pass)
                if m.any():
                    ri = active[m]
                    stack[ri, sp[ri]] = k[m]; sp[ri] += 1
        active = active[sp[active] > 0]
    return t, n_vis, t_tests


def collapse_wide(nodes, K):
    """Greedy BVH2 -> K-wide collapse (the rule of lbvh.cu: bvh4_collapse_host: keep replacing the inner child of largest area by its two
    children until K slots are filled).  Returns (child [n, K] with EMPTY = INT32_MIN, lo [n, K, 3], hi [n, K, 3]), breadth-first."""
    EMPTY = -(1 << 31)
    c = np.stack([nodes["c0"], nodes["c1"]], 1).astype(np.int64)
    lo = np.stack([nodes["lo0"], nodes["lo1"]], 1); hi = np.stack([nodes["hi0"], nodes["hi1"]], 1)
    ext = hi - lo; area = ext[..., 0] * ext[..., 1] + ext[..., 1] * ext[..., 2] + ext[..., 2] * ext[..., 0]
    queue = [0]; out_c, out_src = [], []
    i = 0
    while i < len(queue):
        n2 = queue[i]; i += 1
        slots = [(n2, 0), (n2, 1)]
        while len(slots) < K:
            best, ba = -1, -1.0
            for k, (n, s) in enumerate(slots):
                if c[n, s] >= 0 and area[n, s] > ba: ba, best = area[n, s], k
            if best < 0: break
            n, s_ = slots[best]; ch = int(c[n, s_])
            slots[best] = (ch, 0); slots.append((ch, 1))
        cc = []
        for n, s_ in slots:
            ch = int(c[n, s_])
            if ch >= 0: cc.append(len(queue)); queue.append(ch)
            else: cc.append(ch)
        out_c.append(cc + [EMPTY] * (K - len(slots))); out_src.append(slots + [(0, 0)] * (K - len(slots)))
    child = np.array(out_c, np.int64); src = np.array(out_src, np.int64)
    wlo = lo[src[..., 0], src[..., 1]].astype(np.float64); whi = hi[src[..., 0], src[..., 1]].astype(np.float64)
    empty = child == EMPTY
    wlo[empty] = 3.0e38; whi[empty] = -3.0e38
    return child, wlo, whi


def traverse_wide(wide, tris, O, D):
    """K-wide walk: pop a node, slab-test its K children, visit the hit children nearest first (leaves are tested when reached, with the
    interval re-checked against the current closest hit; inner children go on the stack, farthest first)."""
    child, wlo, whi = wide
    K = child.shape[1]; EMPTY = -(1 << 31)
    R = len(O); inv = 1.0 / np.where(np.abs(D) > 1e-20, D, 1e-20)
    t = np.full(R, np.inf); n_vis = np.zeros(R, np.int64); t_tests = np.zeros(R, np.int64)
    stack = np.zeros((R, 64 * K // 2), np.int64); sp = np.ones(R, np.int64)
    A, E1, E2 = tris[:, 0].astype(np.float64), (tris[:, 1] - tris[:, 0]).astype(np.float64), (tris[:, 2] - tris[:, 0]).astype(np.float64)

    def leaf(rays, ref):
        r = ~ref; first = r >> 2; cnt = (r & 3) + 1
        for j in range(4):
            m = cnt > j
            if not m.any(): break
            ri = rays[m]; ti = first[m] + j
            t_tests[ri] += 1
            o, d = O[ri], D[ri]
            p = np.cross(d, E2[ti]); det = (E1[ti] * p).sum(1); ok = np.abs(det) > 1e-30; idet = 1.0 / np.where(ok, det, 1.0)
            sv = o - A[ti]; u = (sv * p).sum(1) * idet; q = np.cross(sv, E1[ti]); v = (d * q).sum(1) * idet; tt = (E2[ti] * q).sum(1) * idet
            hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (tt > 1e-6) & (tt < t[ri])
            t[ri[hit]] = tt[hit]

    active = np.arange(R)
    while len(active):
        sp[active] -= 1
        n = stack[active, sp[active]]
        n_vis[active] += 1
        o, iv = O[active][:, None, :], inv[active][:, None, :]
        a = (wlo[n] - o) * iv; b = (whi[n] - o) * iv
        tn = np.maximum(np.minimum(a, b).max(2), 0.0); tf = np.maximum(a, b).min(2)
        ch = child[n]
        hit = (tn <= np.minimum(tf, t[active][:, None])) & (ch != EMPTY)
        key = np.where(hit, tn, np.inf)
        order = np.argsort(key, axis=1, kind="stable")
        for j in range(K):                                               # leaves, nearest first
            col = order[:, j]; rows = np.arange(len(active))
            m = hit[rows, col] & (ch[rows, col] < 0) & (key[rows, col] <= t[active])
            if m.any(): leaf(active[m], ch[rows[m], col[m]])
        for j in range(K - 1, -1, -1):                                   # inner children, farthest pushed first
            col = order[:, j]; rows = np.arange(len(active))
            m = hit[rows, col] & (ch[rows, col] >= 0)
            if m.any():
                ri = active[m]; stack[ri, sp[ri]] = ch[rows[m], col[m]]; sp[ri] += 1
        active = active[sp[active] > 0]
    return t, n_vis, t_tests


def quantize_wide(wide, bits=8):
    """Parent-relative quantisation of the child boxes (CWBVH-style): per node an origin and a power-of-two scale per axis, child planes
    rounded outwards to `bits` bits.  Models what a compressed wide node would cost in extra visits."""
    child, wlo, whi = wide
    EMPTY = -(1 << 31); used = child != EMPTY
    lo = np.where(used[..., None], wlo, np.inf).min(1); hi = np.where(used[..., None], whi, -np.inf).max(1)
    qmax = (1 << bits) - 1
    e = np.ceil(np.log2(np.maximum(hi - lo, 1e-30) / qmax)); scale = np.exp2(e)[:, None, :]
    qlo = np.floor((wlo - lo[:, None, :]) / scale); qhi = np.ceil((whi - lo[:, None, :]) / scale)
    dlo = lo[:, None, :] + np.clip(qlo, 0, qmax) * scale; dhi = lo[:, None, :] + np.clip(qhi, 0, qmax) * scale
    dlo[~used] = 3.0e38; dhi[~used] = -3.0e38
    assert np.all(dlo[used] <= wlo[used]) and np.all(dhi[used] >= whi[used])
    return child, dlo, dhi


def make_rays(sc, n_rays, rng):
    vi, pi = B.camera_from_view(sc["camera_view"], sc["aspect"])
    vi = np.asarray(vi, np.float64).reshape(4, 4).T; pi = np.asarray(pi, np.float64).reshape(4, 4).T
    uv = rng.random((n_rays, 2)) * 2 - 1
    tgt = (pi @ np.concatenate([uv, np.ones((n_rays, 1)), np.ones((n_rays, 1))], 1).T).T
    d = tgt[:, :3] / np.linalg.norm(tgt[:, :3], axis=1, keepdims=True)
    D = (vi[:3, :3] @ d.T).T; O = np.broadcast_to(vi[:3, 3], D.shape).copy()
    return O, D


def bounce_rays(O, D, t, rng):
    hit = np.isfinite(t)
    P = O[hit] + D[hit] * t[hit, None]
    # uniform directions on the sphere, pushed off the surface along themselves: the incoherent distribution of a diffuse interior
    w = rng.normal(size=P.shape); w /= np.linalg.norm(w, axis=1, keepdims=True)
    return P + 1e-3 * w, w


def stats(name, n_vis, t_tests):
    return {"set": name, "rays": int(len(n_vis)), "nodes_mean": float(n_vis.mean()), "nodes_p99": float(np.percentile(n_vis, 99)), "nodes_max": int(n_vis.max()),
            "tris_mean": float(t_tests.mean()), "tris_p99": float(np.percentile(t_tests, 99))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene"); ap.add_argument("--rays", type=int, default=20000); ap.add_argument("--out", default=None); ap.add_argument("--seed", type=int, default=1); ap.add_argument("--table", action="store_true"); ap.add_argument("--reinsert", type=float, nargs=2, action="append", default=[], metavar=("PASSES", "FRACTION"), help="also refine the full SAH tree by re-insertion"); ap.add_argument("--quantize", type=int, default=0, help="also model the wide trees with child boxes quantised to this many bits"); ap.add_argument("--wide", type=int, nargs="*", default=[], help="also model K-wide collapses of every tree (4, 8)")
    ap.add_argument("--trav-cost", type=float, nargs="*", default=[1.0], help="node-visit price(s) for the full SAH build (mode 2)")
    a = ap.parse_args()
    sc = load_fixture(a.scene)
    rng = np.random.default_rng(a.seed)
    tri = world_triangles(sc)
    t0 = time.time()
    rb, owner = split_refs(tri)
    keys = morton_keys(rb)
    order = np.argsort(keys, kind="stable")
    rb, owner, keys = rb[order], owner[order], keys[order]
    lb = build_lbvh(rb, keys)
    t_lbvh = time.time() - t0
    t0 = time.time()
    sah, depth_sah, (c0, c1) = B.bvh2_sah_rebuild(lb, 0)
    t_sah = time.time() - t0
    rtris = tri[owner]
    res = {"scene": os.path.basename(a.scene), "triangles": int(len(tri)), "references": int(len(rb)), "lbvh_nodes": int(len(lb)), "lbvh_depth": tree_depth(lb),
           "sah_nodes": int(len(sah)), "sah_depth": int(depth_sah), "sah_cost_lbvh": c0, "sah_cost_rebuilt": c1,
           "host_seconds": {"numpy_lbvh_mirror": round(t_lbvh, 2), "sah_rebuild_cxx": round(t_sah, 3)}, "rays": []}
    trees = [("lbvh", lb, rtris), ("sah_inner", sah, rtris)]
    for passes, frac in a.reinsert:                                      # engine levels 4 (inner rebuild + re-insertion) and 5 (re-insertion alone)
        for tag, src in (("sah_inner+ri", sah), ("lbvh+ri", lb)):
            t0 = time.time()
            ri, d_ri, (c_a, c_b) = B.bvh2_reinsert(src, 0, int(passes), frac)
            res[f"{tag}_p{int(passes)}_f{frac:g}"] = {"depth": int(d_ri), "sah_cost": [c_a, c_b], "host_seconds": round(time.time() - t0, 3)}
            trees.append((f"{tag} p{int(passes)} f{frac:g}", ri, rtris))
    for tc in a.trav_cost:
        t0 = time.time()
        full, perm, d_full, c_full = B.bvh2_sah_build(rb.reshape(-1, 6), tc)
        res[f"sah_full_tc{tc:g}"] = {"nodes": int(len(full)), "depth": int(d_full), "sah_cost": c_full, "host_seconds": round(time.time() - t0, 3)}
        trees.append((f"sah_full_tc{tc:g}", full, rtris[perm]))
        for passes, frac in a.reinsert:
            t0 = time.time()
            ri, d_ri, (c_a, c_b) = B.bvh2_reinsert(full, 0, int(passes), frac)
            res[f"reinsert_p{int(passes)}_f{frac:g}"] = {"depth": int(d_ri), "sah_cost": [c_a, c_b], "host_seconds": round(time.time() - t0, 3)}
            trees.append((f"full+ri p{int(passes)} f{frac:g}", ri, rtris[perm]))
    wides = []
    for K in a.wide:
        for name, nodes, tt in trees:
            w = collapse_wide(nodes, K)
            wides.append((f"{name}/bvh{K}", w, tt))
            if a.quantize: wides.append((f"{name}/bvh{K}q{a.quantize}", quantize_wide(w, a.quantize), tt))
    O, D = make_rays(sc, a.rays, rng)
    base = None
    for name, nodes, tt in trees:
        t, nv, kt = traverse(nodes, tt, O, D)
        if base is None: base = t
        assert np.array_equal(np.isfinite(t), np.isfinite(base)) and np.allclose(t[np.isfinite(t)], base[np.isfinite(base)], rtol=1e-9, atol=1e-12), "all trees must return identical hits"
        res["rays"].append(dict(stats("camera", nv, kt), tree=name))
    for name, w, tt in wides:
        t, nv, kt = traverse_wide(w, tt, O, D)
        assert np.array_equal(np.isfinite(t), np.isfinite(base)) and np.allclose(t[np.isfinite(t)], base[np.isfinite(base)], rtol=1e-9, atol=1e-12)
        res["rays"].append(dict(stats("camera", nv, kt), tree=name))
    O2, D2 = bounce_rays(O, D, base, rng)
    base2 = None
    for name, nodes, tt in trees:
        t, nv, kt = traverse(nodes, tt, O2, D2)
        if base2 is None: base2 = t
        assert np.array_equal(np.isfinite(t), np.isfinite(base2)) and np.allclose(t[np.isfinite(t)], base2[np.isfinite(base2)], rtol=1e-9, atol=1e-12)
        res["rays"].append(dict(stats("bounce", nv, kt), tree=name))
    for name, w, tt in wides:
        t, nv, kt = traverse_wide(w, tt, O2, D2)
        assert np.array_equal(np.isfinite(t), np.isfinite(base2)) and np.allclose(t[np.isfinite(t)], base2[np.isfinite(base2)], rtol=1e-9, atol=1e-12)
        res["rays"].append(dict(stats("bounce", nv, kt), tree=name))
    if a.table:
        for r in res["rays"]: print(f'{r["set"]:7s} {r["tree"]:24s} nodes {r["nodes_mean"]:6.2f} (p99 {r["nodes_p99"]:5.0f})  tris {r["tris_mean"]:5.2f} (p99 {r["tris_p99"]:4.0f})')
    else:
        print(json.dumps(res, indent=1))
    if a.out:
        with open(a.out, "w") as f: json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
