// oracle/shim/CameraCalibration.h — TEST INFRASTRUCTURE.
// src/BundleAdjustment/BundleAdjustment.h:20 includes "CameraCalibration.h" only to see
// MAX_NUMBER_OF_CAMERA_PARAMETERS (defined at src/CalibrationData/CalibrationData.h:19). The real
// header pulls in OpenCV/COLMAP/Boost, which this image lacks; this stand-in shadows it on the include path.
#pragma once
#include <type_traits>
#define MAX_NUMBER_OF_CAMERA_PARAMETERS 17
