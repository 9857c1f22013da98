#!/usr/bin/env python3
"""ORACLE build step (test infrastructure): cut single function definitions out of reference translation units that cannot be
compiled as a whole here (Frame.cc and ORBmatcher.cc need Eigen / Sophus / boost / Pangolin through their headers), so that
oracle/ref_cut_tu.cpp can compile the reference's OWN text of those functions inside minimal stand-in classes.

The pieces are written to oracle/_ref/cut/*.inc -- a git-ignored build directory that the Makefile removes again once the
library is linked; nothing of the reference is stored in the repository.  A piece is located by the start of its definition and ends where its braces balance; the script fails loudly when
a definition is not found exactly once.   usage: cut_reference.py <reference orb_slam3 dir> <output dir>"""
import re
import sys
from pathlib import Path

# (output name, file, regex of the first line of the piece, kind)   kind: "function" = up to the balancing brace, "line" = that line only
PIECES = [
    ("ORBmatcher_TH_HIGH", "src/ORBmatcher.cc", r"^\s*const int ORBmatcher::TH_HIGH\s*=", "line"),
    ("ORBmatcher_TH_LOW", "src/ORBmatcher.cc", r"^\s*const int ORBmatcher::TH_LOW\s*=", "line"),
    ("ORBmatcher_HISTO_LENGTH", "src/ORBmatcher.cc", r"^\s*const int ORBmatcher::HISTO_LENGTH\s*=", "line"),
    ("ORBmatcher_DescriptorDistance", "src/ORBmatcher.cc", r"^\s*int ORBmatcher::DescriptorDistance\(", "function"),
    ("ORBmatcher_ComputeThreeMaxima", "src/ORBmatcher.cc", r"^\s*void ORBmatcher::ComputeThreeMaxima\(", "function"),
    ("Frame_ComputeStereoMatches", "src/Frame.cc", r"^void Frame::ComputeStereoMatches\(\)", "function"),
    ("Frame_ComputeStereoFromRGBD", "src/Frame.cc", r"^void Frame::ComputeStereoFromRGBD\(", "function"),
    ("Frame_AssignFeaturesToGrid", "src/Frame.cc", r"^void Frame::AssignFeaturesToGrid\(\)", "function"),
    ("Frame_PosInGrid", "src/Frame.cc", r"^bool Frame::PosInGrid\(", "function"),
    ("Frame_GetFeaturesInArea", "src/Frame.cc", r"^vector<size_t> Frame::GetFeaturesInArea\(", "function"),
    ("ORBmatcher_ctor", "src/ORBmatcher.cc", r"^\s*ORBmatcher::ORBmatcher\(float nnratio, bool checkOri\)", "function"),
    ("ORBmatcher_RadiusByViewingCos", "src/ORBmatcher.cc", r"^\s*float ORBmatcher::RadiusByViewingCos\(", "function"),
    ("ORBmatcher_SearchByProjection_local", "src/ORBmatcher.cc",
     r"^\s*int ORBmatcher::SearchByProjection\(Frame &F, const vector<MapPoint\*> &vpMapPoints, const float th, const bool bFarPoints", "function"),
    ("ORBmatcher_SearchByBoW_KF_F", "src/ORBmatcher.cc",
     r"^\s*int ORBmatcher::SearchByBoW\(KeyFrame\* pKF,Frame &F, vector<MapPoint\*> &vpMapPointMatches\)", "function"),
    ("ORBmatcher_SearchByProjection_motion", "src/ORBmatcher.cc",
     r"^\s*int ORBmatcher::SearchByProjection\(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono\)", "function"),
    ("ORBmatcher_SearchForInitialization", "src/ORBmatcher.cc",
     r"^\s*int ORBmatcher::SearchForInitialization\(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12, int windowSize\)", "function"),
    ("ORBmatcher_SearchByBoW_KF_KF", "src/ORBmatcher.cc",
     r"^\s*int ORBmatcher::SearchByBoW\(KeyFrame \*pKF1, KeyFrame \*pKF2, vector<MapPoint \*> &vpMatches12\)", "function"),
    ("ORBmatcher_SearchByProjection_reloc", "src/ORBmatcher.cc",
     r"^\s*int ORBmatcher::SearchByProjection\(Frame &CurrentFrame, KeyFrame \*pKF, const set<MapPoint\*> &sAlreadyFound, const float th , const int ORBdist\)", "function"),
    ("ORBmatcher_SearchByProjection_sim3", "src/ORBmatcher.cc",
     r"^\s*int ORBmatcher::SearchByProjection\(KeyFrame\* pKF, Sophus::Sim3f &Scw, const vector<MapPoint\*> &vpPoints,\s*$", "function"),
    ("ORBmatcher_SearchByProjection_sim3_kfs", "src/ORBmatcher.cc",
     r"^\s*int ORBmatcher::SearchByProjection\(KeyFrame\* pKF, Sophus::Sim3<float> &Scw, const std::vector<MapPoint\*> &vpPoints, const std::vector<KeyFrame\*> &vpPointsKFs,\s*$", "function"),
    ("ORBmatcher_Fuse_sim3", "src/ORBmatcher.cc",
     r"^\s*int ORBmatcher::Fuse\(KeyFrame \*pKF, Sophus::Sim3f &Scw, const vector<MapPoint \*> &vpPoints, float th, vector<MapPoint \*> &vpReplacePoint\)", "function"),
    ("ORBmatcher_Fuse_kf", "src/ORBmatcher.cc",
     r"^\s*int ORBmatcher::Fuse\(KeyFrame \*pKF, const vector<MapPoint \*> &vpMapPoints, const float th, const bool bRight\)", "function"),
    ("ORBmatcher_SearchForTriangulation", "src/ORBmatcher.cc", r"^\s*int ORBmatcher::SearchForTriangulation\(KeyFrame \*pKF1, KeyFrame \*pKF2,\s*$", "function"),
    ("KeyFrame_GetFeaturesInArea", "src/KeyFrame.cc", r"^vector<size_t> KeyFrame::GetFeaturesInArea\(", "function"),
    ("KeyFrame_IsInImage", "src/KeyFrame.cc", r"^bool KeyFrame::IsInImage\(", "function"),
    ("MapPoint_PredictScale_KeyFrame", "src/MapPoint.cc", r"^int MapPoint::PredictScale\(const float &currentDist, KeyFrame\* pKF\)", "function"),
    ("MapPoint_PredictScale_Frame", "src/MapPoint.cc", r"^int MapPoint::PredictScale\(const float &currentDist, Frame\* pF\)", "function"),
    ("MapPoint_GetMinDistanceInvariance", "src/MapPoint.cc", r"^float MapPoint::GetMinDistanceInvariance\(\)", "function"),
    ("MapPoint_GetMaxDistanceInvariance", "src/MapPoint.cc", r"^float MapPoint::GetMaxDistanceInvariance\(\)", "function"),
    ("MapPoint_ComputeDistinctiveDescriptors", "src/MapPoint.cc", r"^void MapPoint::ComputeDistinctiveDescriptors\(\)", "function"),
]


def cut(lines, pattern, kind, where):
    hits = [i for i, l in enumerate(lines) if re.search(pattern, l)]
    if len(hits) != 1:
        raise SystemExit(f"cut_reference: {where}: expected exactly one match of {pattern!r}, found {len(hits)}")
    i = hits[0]
    if kind == "line":
        return lines[i]
    depth, seen, out = 0, False, []
    for l in lines[i:]:
        out.append(l)
        code = re.sub(r"//.*", "", l)                      # (none of the pieces has braces in strings or block comments)
        depth += code.count("{") - code.count("}")
        seen = seen or "{" in code
        if seen and depth == 0:
            return "".join(out)
    raise SystemExit(f"cut_reference: {where}: unbalanced braces after {pattern!r}")


def main():
    ref, out = Path(sys.argv[1]), Path(sys.argv[2])
    out.mkdir(parents=True, exist_ok=True)
    cache = {}
    for name, rel, pattern, kind in PIECES:
        lines = cache.setdefault(rel, (ref / rel).read_text(errors="replace").splitlines(keepends=True))
        (out / f"{name}.inc").write_text(cut(lines, pattern, kind, f"{rel}:{name}"))


if __name__ == "__main__":
    main()
