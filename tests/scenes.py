"""Seeded scenes for the matcher tests (shared by the CPU pinning tests and the GPU parity tests): the flat arrays that both the
reference-cut glue (oracle/ref.py) and the GPU host adapter harness (tests/host/matcher_host.cpp) take."""
import numpy as np

from oracle import port
from orb_slam3_ros_b200 import synth


def local_points_scene(with_stereo, seed=31, nmp=1600):
    """Tracking::SearchLocalPoints: more projected map points than key points, so that many compete for the same key point"""
    h, w = 480, 752
    pe = port.PortExtractor(1000, 1.2, 8)
    _, k, d, _ = pe.extract(synth.frame(h, w, 8))
    n = len(k)
    rng = np.random.default_rng(seed)
    grid4 = np.float32([0.0, 0.0, np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)])
    src = rng.integers(0, n, nmp)
    proj = np.stack([k["x"][src] + rng.normal(0, 2.5, nmp), k["y"][src] + rng.normal(0, 2.5, nmp),
                     k["x"][src] - rng.uniform(0, 40, nmp), rng.choice([0.9, 0.999, 0.9985], nmp)], 1).astype(np.float32)
    level = np.clip(k["octave"][src] + rng.integers(-1, 2, nmp), 0, 7).astype(np.int32)
    mp_desc = d[src].copy()
    mp_desc[:, :4] ^= rng.integers(0, 256, (nmp, 4), dtype=np.uint8) & rng.integers(0, 256, (nmp, 4), dtype=np.uint8)
    in_view = (rng.random(nmp) < 0.9).astype(np.uint8)
    has_point = (rng.random(n) < 0.25).astype(np.uint8)
    u_right = np.where(rng.random(n) < 0.5, k["x"] - rng.uniform(1, 40, n), -1).astype(np.float32) if with_stereo else None
    return dict(k=k, d=d, grid4=grid4, sf=pe.scale_factors, proj=proj, level=level, mp_desc=mp_desc, in_view=in_view, has_point=has_point,
                u_right=u_right)


def motion_scene(stereo, direction=0, seed=5, dense=False):
    """Tracking::TrackWithMotionModel: the last frame's map points projected into the current frame with the predicted pose.
    direction: 0 = sideways motion (level window oct-1 .. oct+1), +1 / -1 = forward / backward by more than the baseline (stereo
    only: levels >= oct / <= oct).  A fraction of the last frame's points are temporal stereo points (no observations) and some
    current key points already hold points -- both kinds of in-loop state of ORBmatcher.cc:1749-1751.  dense: several last-frame
    points land on the same key points (collisions)."""
    h, w = 480, 752
    rng = np.random.default_rng(seed)
    seq = synth.sequence(h, w, 40, canvas=1024, base_seed=900 + seed)
    pe = port.PortExtractor(1000, 1.2, 8)
    _, kl, dl, _ = pe.extract(seq[10])
    _, kc, dc, _ = pe.extract(seq[11])
    # image motion between the two crops (see synth.sequence): the texture moves by (-dx, -dy)
    n_seq, span_x, span_y = 40, 1024 - w, 1024 - h
    off = lambda i: (int(round((0.5 + 0.5 * np.sin(2 * np.pi * i / (n_seq - 1))) * span_x)), int(round((0.5 + 0.5 * np.cos(np.pi * i / (n_seq - 1))) * span_y)))
    (x0, y0), (x1, y1) = off(10), off(11)
    dx, dy = float(x0 - x1), float(y0 - y1)
    fx = fy = np.float32(458.0)
    cx, cy = np.float32(w / 2), np.float32(h / 2)
    m = len(kl)
    depth = rng.uniform(4.0, 9.0, m).astype(np.float32)
    # last frame at the origin: X = depth * K^-1 (u, v, 1)
    pos = np.stack([(kl["x"] - cx) / fx * depth, (kl["y"] - cy) / fy * depth, depth], 1).astype(np.float32)
    zmean = 6.5
    tz = {0: 0.0, 1: -0.9, -1: 0.9}[direction]                    # camera moves forward: points come closer (z decreases in the camera frame)
    t = np.float32([dx * zmean / float(fx), dy * zmean / float(fy), tz])
    ang = 0.01
    R = np.float32([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    Tcw = np.concatenate([R.ravel(), t]).astype(np.float32)
    Tlw = np.concatenate([np.eye(3, dtype=np.float32).ravel(), np.zeros(3, np.float32)])
    last_state = rng.choice([0, 1, 2], m, p=[0.15, 0.7, 0.15]).astype(np.uint8)
    last_desc = dl.copy()
    last_desc[:, :3] ^= rng.integers(0, 256, (m, 3), dtype=np.uint8) & rng.integers(0, 256, (m, 3), dtype=np.uint8)
    if dense:                                                     # duplicate points: several map points project onto the same place
        dup = rng.integers(0, m, m // 2)
        sel = np.arange(m // 2)
        pos[sel] = pos[dup] + rng.normal(0, 0.01, (m // 2, 3)).astype(np.float32)
        last_desc[sel] = last_desc[dup]
        last_state[sel] = rng.choice([1, 2], m // 2, p=[0.8, 0.2])
    n = len(kc)
    cur = dict(kps_xy=np.stack([kc["x"], kc["y"]], 1), octaves=kc["octave"].astype(np.int32), angles=kc["angle"], desc=dc,
               u_right=(np.where(rng.random(n) < 0.6, kc["x"] - rng.uniform(2, 60, n), -1).astype(np.float32) if stereo else None),
               state=rng.choice([0, 1, 2], n, p=[0.8, 0.12, 0.08]).astype(np.uint8),
               fp=np.float32([0, w, 0, h, np.float32(64) / np.float32(w), np.float32(48) / np.float32(h), 40.0 if stereo else 0.0, 0.11 if stereo else 0.0]),
               scale_factors=pe.scale_factors, Tcw=Tcw, cam4=np.float32([fx, fy, cx, cy]))
    last_angle = kl["angle"].copy()
    turn = rng.random(m) < 0.2
    last_angle[turn] = (last_angle[turn] + rng.uniform(40, 320, int(turn.sum())).astype(np.float32)) % np.float32(360)      # inconsistent rotations
    last = dict(octaves=kl["octave"].astype(np.int32), angles=last_angle, state=last_state, outlier=(rng.random(m) < 0.05).astype(np.uint8),
                pos=pos, desc=last_desc, Tlw=Tlw)
    return cur, last


def init_scene(seed=3, crowd=False, jitter=0.0):
    """Tracking::MonocularInitialization: two consecutive frames of a sequence; the search centres are the first frame's own key points
    (Tracking.cc:2484-2486), optionally jittered.  crowd: groups of level-0 key points of the first frame share one descriptor, so that
    several of them claim the same key point of the second frame -- the take-over / vMatchedDistance paths of ORBmatcher.cc:687, :706-710."""
    h, w = 480, 752
    rng = np.random.default_rng(seed)
    seq = synth.sequence(h, w, 40, canvas=1024, base_seed=700 + seed)
    pe = port.PortExtractor(1000, 1.2, 8)
    _, k1, d1, _ = pe.extract(seq[20])
    _, k2, d2, _ = pe.extract(seq[21])
    d1 = d1.copy()
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    if jitter:
        prev += rng.normal(0, jitter, prev.shape).astype(np.float32)
    if crowd:
        lvl0 = np.flatnonzero(k1["octave"] == 0)
        for g in range(len(lvl0) // 4):                            # groups of four neighbours in index order
            grp = lvl0[4 * g:4 * g + 4]
            d1[grp] = d1[grp[0]]
            d1[grp[1:], 0] ^= rng.integers(0, 4, 3, dtype=np.uint8)      # a bit or two apart: different distances to the same target
            prev[grp] = prev[grp[0]]
    f1 = dict(octaves=k1["octave"].astype(np.int32), angles=k1["angle"].copy(), desc=d1)
    f2 = dict(kps_xy=np.stack([k2["x"], k2["y"]], 1).astype(np.float32), octaves=k2["octave"].astype(np.int32), angles=k2["angle"].copy(), desc=d2,
              fp=np.float32([0, w, 0, h, np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)]))
    return f1, f2, prev


def bow_scene(seed=17, levelsup=2):
    """Tracking::TrackReferenceKeyFrame: a key frame and a frame of the same place (the two views of a stereo pair), their feature vectors
    from the BoW transform; 85 % of the key-frame features hold a map point, a quarter of them with an inconsistent rotation."""
    from orb_slam3_ros_b200.bow import synthetic_vocabulary
    vocab = synthetic_vocabulary(8, 4, seed=5, stop_fraction=0.0)
    left, right = synth.stereo_pair(376, 620, 4, dmax=25)
    _, kk, dk, _ = port.PortExtractor(900, 1.2, 8).extract(left)
    _, kf, df, _ = port.PortExtractor(900, 1.2, 8).extract(right)
    rng = np.random.default_rng(seed)
    ang_k = kk["angle"].copy()
    turn = rng.random(len(kk)) < 0.25
    ang_k[turn] = (ang_k[turn] + rng.uniform(40, 320, int(turn.sum())).astype(np.float32)) % np.float32(360)      # inconsistent rotations
    has_point = (rng.random(len(kk)) < 0.85).astype(np.uint8)
    fv_k = port.bow_transform(vocab, dk, levelsup, 1)[2:5]
    fv_f = port.bow_transform(vocab, df, levelsup, 1)[2:5]
    return dict(ang_k=ang_k, dk=dk, has_point=has_point, fv_k=fv_k, ang_f=kf["angle"].copy(), df=df, fv_f=fv_f, nk=len(kk), nf=len(kf))


def reloc_scene(seed=5, th=10.0):
    """Tracking::Relocalization, refinement step: a candidate key frame's map points projected into the current frame with its (PnP) pose.
    Built on motion_scene's geometry (the key frame takes the place of the last frame): some points are bad, some were found already, some
    lie outside their scale-invariance distance range, some current key points already hold a point, and a dense half of the map points
    projects onto the same places (collisions)."""
    cur, last = motion_scene(False, 0, seed=seed, dense=True)
    rng = np.random.default_rng(100 + seed)
    m, n = len(last["octaves"]), len(cur["octaves"])
    sf = np.asarray(cur["scale_factors"], np.float32)
    depth = last["pos"][:, 2].astype(np.float32)
    # MapPoint::UpdateNormalAndDepth: mfMaxDistance = dist * scaleFactor[level], mfMinDistance = mfMaxDistance / scaleFactor[nLevels - 1]
    max_dist = (depth * sf[last["octaves"]]).astype(np.float32)
    min_dist = (max_dist / sf[-1]).astype(np.float32)
    far = rng.random(m) < 0.1
    max_dist[far] *= np.float32(0.5)                                 # out of range: skipped by the distance test
    state = rng.choice([0, 1, 2, 3], m, p=[0.1, 0.7, 0.1, 0.1]).astype(np.uint8)
    fp = np.concatenate([cur["fp"][:8], np.float32([len(sf), np.log(np.float32(1.2))])]).astype(np.float32)
    c = dict(kps_xy=cur["kps_xy"], octaves=cur["octaves"], angles=cur["angles"], desc=cur["desc"], holds=(rng.random(n) < 0.2).astype(np.uint8), fp=fp,
             scale_factors=sf, Tcw=cur["Tcw"], cam4=cur["cam4"])
    kf = dict(angles=last["angles"], state=state, pos=last["pos"], desc=last["desc"], min_dist=min_dist, max_dist=max_dist)
    return c, kf


def sim3_scene(seed=5, scale=1.07):
    """LoopClosing: the map points of a loop candidate's neighbourhood projected into the current key frame with a Sim3 (scale != 1).  Built on
    reloc_scene: the key frame is its current frame, the map points are its key-frame points expressed in a frame that is `scale` times larger;
    some normals point away (viewing-angle test), some points are bad, some key points are matched already."""
    cur, kf = reloc_scene(seed)
    rng = np.random.default_rng(200 + seed)
    m = len(kf["state"])
    pos = (kf["pos"] * np.float32(scale)).astype(np.float32)                      # world = scale * camera-frame coordinates of the old scene
    R, t = cur["Tcw"][:9].astype(np.float32), cur["Tcw"][9:].astype(np.float32)
    # Scw = (s, R, t_s): the reference uses Tcw = (R, t_s / s).  With world = scale * old world, s = 1 / scale and t_s = t the camera-frame points are
    # scale * the old ones: the same pixels, depths and distances scaled
    sim3 = np.concatenate([R, t, np.float32([1.0 / scale])]).astype(np.float32)
    normal = pos / np.linalg.norm(pos, axis=1, keepdims=True)                   # mean viewing direction: from the (old) camera towards the point
    away = rng.random(m) < 0.15
    normal[away] *= -1
    state = np.where(kf["state"] == 2, 2, 1).astype(np.uint8)
    k = dict(kps_xy=cur["kps_xy"], octaves=cur["octaves"], desc=cur["desc"], held=(rng.random(len(cur["octaves"])) < 0.3).astype(np.uint8),
             fp=cur["fp"], scale_factors=cur["scale_factors"], cam4=cur["cam4"])
    pts = dict(state=state, pos=pos, normal=normal.astype(np.float32), desc=kf["desc"], min_dist=(kf["min_dist"] * np.float32(scale)).astype(np.float32),
               max_dist=(kf["max_dist"] * np.float32(scale)).astype(np.float32))
    return k, pts, sim3


def fuse_scene(seed=5, stereo=False):
    """LocalMapping::SearchInNeighbors: a neighbour's map points fused into a key frame.  Built on reloc_scene (key frame = its current frame with
    its pose); the key points hold good / bad / no map points with different observation counts, so that both replacement directions and new
    observations occur; the list contains null entries, bad points and duplicates; with `stereo` most key points have a right coordinate (the 7.8
    chi-square branch), a few of them inconsistent with the projection."""
    cur, kfp = reloc_scene(seed)
    rng = np.random.default_rng(400 + seed)
    n, m = len(cur["octaves"]), len(kfp["state"])
    sf = np.asarray(cur["scale_factors"], np.float32)
    pos = kfp["pos"]
    normal = (pos / np.linalg.norm(pos, axis=1, keepdims=True)).astype(np.float32)
    fp = cur["fp"].copy()
    fp[6] = np.float32(40.0)                                                     # mbf
    R, t = cur["Tcw"][:9].reshape(3, 3).astype(np.float32), cur["Tcw"][9:].astype(np.float32)
    u_right = None
    if stereo:
        # right coordinates consistent with a plausible depth for most key points (so that the 3-dof test can pass), noise on some
        z = rng.uniform(4.0, 9.0, n).astype(np.float32)
        u_right = (cur["kps_xy"][:, 0] - np.float32(40.0) / z).astype(np.float32)
        u_right[rng.random(n) < 0.3] = -1
    state = rng.choice([0, 1, 2], m, p=[0.05, 0.85, 0.10]).astype(np.uint8)
    dup = rng.integers(0, m, 30)                                                   # (the harness cannot alias entries: duplicates are separate but identical points)
    k = dict(kps_xy=cur["kps_xy"], octaves=cur["octaves"], desc=cur["desc"], held=rng.choice([0, 1, 2], n, p=[0.5, 0.42, 0.08]).astype(np.uint8),
             held_obs=rng.integers(1, 9, n).astype(np.int32), u_right=u_right, inv_sigma2=(1.0 / (sf * sf)).astype(np.float32), fp=fp, scale_factors=sf,
             Tcw=cur["Tcw"], cam4=cur["cam4"])
    pts = dict(state=state, obs=rng.integers(1, 9, m).astype(np.int32), pos=pos, normal=normal, desc=kfp["desc"], min_dist=kfp["min_dist"], max_dist=kfp["max_dist"])
    for a, b in zip(dup[:15], dup[15:]):
        for key in ("pos", "normal", "desc", "min_dist", "max_dist"):
            pts[key][b] = pts[key][a]
    return k, pts


def triangulation_scene(seed=17, levelsup=2, stereo=False, ties=False):
    """LocalMapping::CreateNewMapPoints: two neighbouring key frames (the two views of a stereo pair: a pure sideways translation, so that the
    epipolar lines are the image rows and true matches satisfy the constraint), feature vectors from the BoW transform, most key points
    without map points; with `stereo` some key points have right coordinates (no epipole test for those, and bOnlyStereo has something to keep)."""
    sc = bow_scene(seed=seed, levelsup=levelsup)
    from orb_slam3_ros_b200 import synth as _s
    left, right = _s.stereo_pair(376, 620, 4, dmax=25)
    pe = port.PortExtractor(900, 1.2, 8)
    _, ka, da, _ = pe.extract(left)
    _, kb, db, _ = pe.extract(right)
    rng = np.random.default_rng(500 + seed)
    sf = np.asarray(pe.scale_factors, np.float32)
    eye = np.eye(3, dtype=np.float32).ravel()
    T1 = np.concatenate([eye, np.float32([0, 0, 0])]).astype(np.float32)
    T2 = np.concatenate([eye, np.float32([-0.25, 0.0, 0.02])]).astype(np.float32)      # camera 2 to the right (and slightly forward: a finite epipole)

    def kf(k, d, fv, T):
        n = len(k)
        return dict(kps_xy=np.stack([k["x"], k["y"]], 1).astype(np.float32), octaves=k["octave"].astype(np.int32), angles=k["angle"].copy(), desc=d,
                    has_point=(rng.random(n) < 0.25).astype(np.uint8),
                    u_right=(np.where(rng.random(n) < 0.4, k["x"] - rng.uniform(2, 40, n), -1).astype(np.float32) if stereo else None), fv=fv, Tcw=T)
    if ties:                                                                     # equal distances inside a node: the LAST minimum wins in the reference (:1010)
        db = db.copy()
        node, start, feat = (np.asarray(a) for a in sc["fv_f"])
        ends = list(start[1:]) + [len(feat)]
        for s0, e0 in zip(start, ends):
            for a in range(s0 + 1, e0, 2):
                db[feat[a]] = db[feat[a - 1]]
    k1, k2 = kf(ka, da, sc["fv_k"], T1), kf(kb, db, sc["fv_f"], T2)
    common = dict(sigma2=(sf * sf).astype(np.float32), scale_factors=sf, cam4=np.float32([458.0, 458.0, 310.0, 188.0]))
    return k1, k2, common


def fisheye_local_points_scene(seed=41, nmp=1500):
    """Tracking::SearchLocalPoints on a fisheye-stereo frame (Nleft != -1): the two views of a stereo pair as the two eyes, stereo partners between
    them for a third of the key points (mvLeftToRightMatch / mvRightToLeftMatch), map points visible in the left eye, the right eye or both."""
    left, right = synth.stereo_pair(376, 620, 4, dmax=25)
    pe = port.PortExtractor(900, 1.2, 8)
    _, kl, dl, _ = pe.extract(left)
    _, kr, dr, _ = pe.extract(right)
    nL, nR = len(kl), len(kr)
    rng = np.random.default_rng(seed)
    l2r, r2l = np.full(nL, -1, np.int32), np.full(nR, -1, np.int32)
    perm = rng.permutation(min(nL, nR))[: min(nL, nR) // 3]
    partner = rng.permutation(nR)[: len(perm)]
    for a, b in zip(perm, partner):
        l2r[a] = b
        r2l[b] = a
    w, h = 620, 376
    fp = np.float32([0, w, 0, h, np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)])
    srcL, srcR = rng.integers(0, nL, nmp), rng.integers(0, nR, nmp)
    proj_l = np.stack([kl["x"][srcL] + rng.normal(0, 2.5, nmp), kl["y"][srcL] + rng.normal(0, 2.5, nmp), rng.choice([0.9, 0.999, 0.9985], nmp)], 1).astype(np.float32)
    proj_r = np.stack([kr["x"][srcR] + rng.normal(0, 2.5, nmp), kr["y"][srcR] + rng.normal(0, 2.5, nmp), rng.choice([0.9, 0.999, 0.9985], nmp)], 1).astype(np.float32)
    level_l = np.clip(kl["octave"][srcL] + rng.integers(-1, 2, nmp), 0, 7).astype(np.int32)
    level_r = np.clip(kr["octave"][srcR] + rng.integers(-1, 2, nmp), 0, 7).astype(np.int32)
    level_r[rng.random(nmp) < 0.1] = -1
    in_l = (rng.random(nmp) < 0.7).astype(np.uint8)
    in_r = (rng.random(nmp) < 0.6).astype(np.uint8)
    use_r = rng.random(nmp) < 0.4                                                  # descriptor close to the right-eye key point for some
    desc = np.where(use_r[:, None], dr[srcR], dl[srcL]).astype(np.uint8)
    desc[:, :4] ^= rng.integers(0, 256, (nmp, 4), dtype=np.uint8) & rng.integers(0, 256, (nmp, 4), dtype=np.uint8)
    f = dict(kps_l=np.stack([kl["x"], kl["y"]], 1).astype(np.float32), oct_l=kl["octave"].astype(np.int32),
             kps_r=np.stack([kr["x"], kr["y"]], 1).astype(np.float32), oct_r=kr["octave"].astype(np.int32), desc=np.concatenate([dl, dr]), fp=fp, l2r=l2r, r2l=r2l,
             has_point=(rng.random(nL + nR) < 0.2).astype(np.uint8), scale_factors=pe.scale_factors)
    mp = dict(proj_l=proj_l, level_l=level_l, in_view_l=in_l, proj_r=proj_r, level_r=level_r, in_view_r=in_r, desc=desc)
    return f, mp


def fisheye_motion_scene(direction=0, seed=5, dense=False):
    """Tracking::TrackWithMotionModel on a fisheye-stereo rig: motion_scene's current frame as the LEFT eye; the right eye sees the same key points
    through GetRelativePoseTrl() (a sideways baseline), re-projected with the left camera model as the reference does, with its own jitter and
    order; the last frame's features are split into a left and a right set."""
    cur, last = motion_scene(True, direction, seed=seed, dense=dense)
    rng = np.random.default_rng(600 + seed)
    nL = len(cur["octaves"])
    fx, fy, cx, cy = (np.float32(v) for v in cur["cam4"])
    Trl = np.concatenate([np.eye(3, dtype=np.float32).ravel(), np.float32([-0.11, 0.0, 0.0])]).astype(np.float32)
    # right-eye key points: the left ones shifted by the disparity of a plausible depth, shuffled, some dropped
    z = rng.uniform(4.0, 9.0, nL).astype(np.float32)
    keep = np.flatnonzero(rng.random(nL) < 0.85)
    rng.shuffle(keep)
    kps_r = np.stack([cur["kps_xy"][keep, 0] - fx * np.float32(0.11) / z[keep] + rng.normal(0, 0.3, len(keep)).astype(np.float32),
                      cur["kps_xy"][keep, 1] + rng.normal(0, 0.3, len(keep)).astype(np.float32)], 1).astype(np.float32)
    inside = (kps_r[:, 0] > 1) & (kps_r[:, 0] < 750) & (kps_r[:, 1] > 1) & (kps_r[:, 1] < 478)
    keep, kps_r = keep[inside], kps_r[inside]
    desc_r = cur["desc"][keep].copy()
    desc_r[:, :2] ^= rng.integers(0, 256, (len(keep), 2), dtype=np.uint8) & rng.integers(0, 256, (len(keep), 2), dtype=np.uint8)
    nR = len(keep)
    c = dict(kps_l=cur["kps_xy"], oct_l=cur["octaves"], ang_l=cur["angles"], kps_r=kps_r, oct_r=cur["octaves"][keep], ang_r=cur["angles"][keep],
             desc=np.concatenate([cur["desc"], desc_r]), state=np.concatenate([cur["state"], rng.choice([0, 1, 2], nR, p=[0.8, 0.12, 0.08]).astype(np.uint8)]),
             fp=cur["fp"], scale_factors=cur["scale_factors"], Tcw=cur["Tcw"], Trl=Trl, cam4=cur["cam4"])
    m = len(last["octaves"])
    lst = dict(last, n_left=int(m * 0.55))
    return c, lst
