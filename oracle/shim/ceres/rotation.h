// oracle/shim/ceres/rotation.h — TEST INFRASTRUCTURE. The reference includes <ceres/rotation.h>
// (src/BundleAdjustment/BundleAdjustment.h:15) but uses nothing from it.
#pragma once
