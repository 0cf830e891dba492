// TEST INFRASTRUCTURE — separate TU because KDOPBroadPhase.h and AABBBroadPhase.h both define NodeComparator.
#include "AABBBroadPhase.h"
BroadPhase *ref_make_aabb() { return new AABBBroadPhase(); }
