// TEST INFRASTRUCTURE — separate TU because KDOPBroadPhase.h and AABBBroadPhase.h both define NodeComparator.
#include "KDOPBroadPhase.h"
BroadPhase *ref_make_kdop() { return new KDOPBroadPhase(); }
