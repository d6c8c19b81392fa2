// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// Pre-included (-include) when oracle/ref_build.sh compiles the reference's sources into oracle/_ref/TestRenderStereoPanorama_b200:
// it applies the two-line factory change of integration/reference_tree/optical_flow/B200Flow.h WITHOUT editing the reference's
// OpticalFlowFactory.h (which lies read-only under /root/reference): the reference's factory is compiled under another name and a
// factory with the two extra names forwards everything else to it.
#pragma once
#include "NovelView.h"   // PixFlow.h relies on the using-directives of the headers the reference includes before it
#define makeOpticalFlowByName makeOpticalFlowByName_reference
#include "OpticalFlowFactory.h"
#undef makeOpticalFlowByName
#include "B200Flow.h"

namespace surround360 {
namespace optical_flow {
static OpticalFlowInterface* makeOpticalFlowByName(const string flowAlgName) {
  if (flowAlgName == "b200_pixflow_low") return new B200Flow("pixflow_low");
  if (flowAlgName == "b200_pixflow_search_20") return new B200Flow("pixflow_search_20");
  return makeOpticalFlowByName_reference(flowAlgName);
}
} // namespace optical_flow
} // namespace surround360
