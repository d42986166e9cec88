// oracle/shim/glog/logging.h — TEST INFRASTRUCTURE. The reference includes <glog/logging.h>
// (src/BundleAdjustment/BundleAdjustment.h:16) but uses nothing from it on this path.
#pragma once
