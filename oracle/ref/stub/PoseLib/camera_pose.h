// ORACLE — TEST INFRASTRUCTURE ONLY.
// Shadows PoseLib/camera_pose.h when the reference's robust/ransac_impl.h is compiled for oracle/_ref: that header only
// needs the NAME CameraPose (default template argument of ransac<> / score_models<>); the real class is Eigen code.
#pragma once
namespace poselib {
struct CameraPose;
}
