// TEST INFRASTRUCTURE.  The reference's ORBmatcher.cc compiled where it lies under another class name, so that it can live in one
// binary next to the drop-in shell that implements ORB_SLAM2::ORBmatcher on the GPU (tests/cpp/ref_twin_test.cc).  Nothing is
// copied: the reference translation unit is pulled in by #include through -I$(REF)/src.
#define ORBmatcher ORBmatcherRef
#include "ORBmatcher.cc"
