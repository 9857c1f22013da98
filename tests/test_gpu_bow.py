"""GPU parity of the bag-of-words transform ("next" row: Frame::ComputeBoW, Frame.cc:738-745 ->
DBoW2 TemplatedVocabulary::transform, TemplatedVocabulary.h:1139-1275) against the std::map-based oracle and -- when
oracle/_ref/liborbref.so is present -- against the reference's vendored DBoW2 itself (loadFromTextFile + transform), on seeded
synthetic vocabulary trees (ORBvoc.txt.bin is not available offline)."""
import tempfile
from pathlib import Path

import numpy as np
import pytest

from oracle import port, ref
from orb_slam3_ros_b200.bow import Vocabulary, synthetic_vocabulary

pytestmark = pytest.mark.gpu


def _descriptors(vocab, n, rng, near=0.7):
    """descriptors near random nodes (so sets share words) mixed with uniform random ones"""
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    leaves = np.flatnonzero(vocab["child_count"] == 0)
    pick = rng.random(n) < near
    src = vocab["node_desc"][rng.choice(leaves[: max(8, len(leaves) // 20)], n)]
    bits = np.unpackbits(src, axis=1)
    flips = rng.integers(0, 256, (n, 12))
    np.bitwise_xor.at(bits, (np.repeat(np.arange(n), 12), flips.ravel()), 1)
    d[pick] = np.packbits(bits, axis=1)[pick]
    return d


_dbow2 = {}


def _vendored_dbow2(vocab):
    """the same tree loaded by the reference's DBoW2 (its text loader accepts k <= 20, L <= 10); None when not available"""
    if not ref.available() or int(vocab["child_count"].max()) > 20 or vocab["depth"] > 10:
        return None
    key = (id(vocab), vocab["node_desc"].tobytes(), vocab["node_weight"].tobytes())
    if key not in _dbow2:
        with tempfile.TemporaryDirectory() as d:
            ref.write_vocabulary_text(vocab, Path(d) / "voc.txt")
            _dbow2.clear()
            _dbow2[key] = ref.RefVocabulary(Path(d) / "voc.txt")
    return _dbow2[key]


def _check(vocab, sets, levelsup, norm):
    v = Vocabulary(vocab)
    got = v.transform(sets, levelsup, norm)
    rv = _vendored_dbow2(vocab) if norm == 1 else None      # ORBvoc uses L1 scoring (header `.. 0 0`)
    if rv is not None:
        for s, g in zip(sets, got):
            w = rv.transform(s, levelsup)
            assert all(np.array_equal(a, b) for a, b in zip(w[:5], g[:5])) and w[5] == g[5], "vs the vendored DBoW2"
    for s, g in zip(sets, got):
        w = port.bow_transform(vocab, s, levelsup, norm)
        assert np.array_equal(w[0], g[0])
        assert np.array_equal(w[1], g[1])                  # doubles bit-exact: same addition order as the std::map
        assert np.array_equal(w[2], g[2]) and np.array_equal(w[3], g[3]) and np.array_equal(w[4], g[4])
        assert w[5] == g[5]
    v.close()
    return got


@pytest.mark.parametrize("k,depth,ragged", [(10, 3, False), (10, 4, False), (6, 5, True), (40, 2, False)])
def test_bow_transform_matches_oracle(k, depth, ragged):
    vocab = synthetic_vocabulary(k, depth, seed=3 + k, ragged=ragged)
    rng = np.random.default_rng(11)
    sets = [_descriptors(vocab, n, rng) for n in (1000, 1, 0, 257, 2003)]
    for levelsup in (4, 2, 0, depth + 3):                  # depth - levelsup <= 0 -> every feature under the root (:1240)
        for norm in (1, 2, 0):
            got = _check(vocab, sets, levelsup, norm)
    assert got[0][5] > 0 and len(got[0][0]) < got[0][5]    # words really shared between features
    assert got[2][5] == 0 and len(got[2][0]) == 0          # empty set


def test_bow_ties_take_first_child():
    """identical sibling descriptors: strict '<' keeps the first child in visiting order (TemplatedVocabulary.h:1256)"""
    vocab = synthetic_vocabulary(8, 3, seed=2)
    for p in range(len(vocab["child_begin"])):
        b, c = vocab["child_begin"][p], vocab["child_count"][p]
        if c:
            ids = vocab["child_list"][b:b + c]
            vocab["node_desc"][ids[1::2]] = vocab["node_desc"][ids[0::2]][: len(ids[1::2])]
    rng = np.random.default_rng(0)
    _check(vocab, [_descriptors(vocab, 500, rng)], 2, 1)


def test_bow_all_stopped():
    vocab = synthetic_vocabulary(5, 2, seed=1)
    vocab["node_weight"][:] = 0.0
    rng = np.random.default_rng(0)
    got = _check(vocab, [rng.integers(0, 256, (64, 32), dtype=np.uint8)], 1, 1)
    assert got[0][5] == 0
