"""Constants of the reference's matcher (orb_slam3/src/ORBmatcher.cc:35-37) and extractor (ORBextractor.cc:71-73)."""
TH_HIGH = 100
TH_LOW = 50
HISTO_LENGTH = 30
PATCH_SIZE = 31
HALF_PATCH_SIZE = 15
EDGE_THRESHOLD = 19
