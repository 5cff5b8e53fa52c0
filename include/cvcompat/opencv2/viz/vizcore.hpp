// cvcompat/opencv2/viz/vizcore.hpp -- HEADLESS stand-in for the handful of cv::viz names the reference's apps/demo.cpp uses
// (apps/demo.cpp:10-33,63-67,106-128,135), so that the application compiles and links UNCHANGED against include/kfusion +
// libkfusion.so on a machine without OpenCV or a display.  Widgets are empty types, the viewer remembers its pose and never
// asks to stop; nothing is drawn.  With a real OpenCV on the include path this directory is simply not used.
#pragma once
#include <opencv2/core/core.hpp>

namespace cv { namespace viz {

struct KeyboardEvent
{
    enum Action { KEY_UP = 0, KEY_DOWN = 1 };
    Action action;
    unsigned char code;
    KeyboardEvent() : action(KEY_UP), code(0) {}
};

struct Color
{
    double b, g, r;
    Color(double b_ = 0, double g_ = 0, double r_ = 0) : b(b_), g(g_), r(r_) {}
    static Color apricot() { return Color(177, 206, 251); }
    static Color white() { return Color(255, 255, 255); }
};

struct Widget {};
struct Widget3D : Widget {};
struct WCube : Widget3D { WCube(const Vec3d & = Vec3d::all(-0.5), const Vec3d & = Vec3d::all(0.5), bool = true, const Color & = Color::white()) {} };
struct WCoordinateSystem : Widget3D { explicit WCoordinateSystem(double = 1.0) {} };
struct WCloud : Widget3D { explicit WCloud(const Mat &, const Color & = Color::white()) {} };

class Viz3d
{
public:
    typedef void (*KeyboardCallback)(const KeyboardEvent &, void *);
    explicit Viz3d(const String & = String()) : cb_(0), cookie_(0) {}
    void showWidget(const String &, const Widget &, const Affine3f & = Affine3f::Identity()) {}
    void registerKeyboardCallback(KeyboardCallback cb, void *cookie = 0) { cb_ = cb; cookie_ = cookie; }
    Affine3f getViewerPose() const { return pose_; }
    void setViewerPose(const Affine3f &p) { pose_ = p; }
    bool wasStopped() const { return false; }
    void spinOnce(int = 1, bool = false) {}
private:
    Affine3f pose_;
    KeyboardCallback cb_;
    void *cookie_;
};

}}  // namespace cv::viz
