// tests/stubs/opencv2/core.hpp -- TEST INFRASTRUCTURE.  A minimal stand-in for <opencv2/core.hpp> (OpenCV C++ headers are not
// installed in this image) with just the declarations include/sindyn_classes.hpp touches under SINDYN_WITH_OPENCV, so that
// the reference-signature overloads (DynaDetect.h:98-131, ORBextractor.h:54-64) are at least type-checked and exercised on
// the host.  Semantics follow cv::Mat / cv::_InputArray / cv::_OutputArray for continuous 2-D matrices only.
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_16U 2
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC(n) CV_MAKETYPE(CV_8U, (n))
#define CV_8UC1 CV_8UC(1)
#define CV_8UC3 CV_8UC(3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)

namespace cv {

class Mat {
public:
    unsigned char *data = nullptr;
    int rows = 0, cols = 0;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void *ext) : data((unsigned char *)ext), rows(r), cols(c), type_(type) { step = (size_t)c * elemSize(); }
    void create(int r, int c, int type)
    {
        type_ = type; rows = r; cols = c; step = (size_t)c * elemSize();
        own_ = std::shared_ptr<unsigned char>(new unsigned char[step * (size_t)r + 1], std::default_delete<unsigned char[]>());
        data = own_.get();
    }
    void release() { own_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize1() const { return CV_MAT_DEPTH(type_) == CV_16U ? 2 : 1; }
    size_t elemSize() const { return elemSize1() * (size_t)channels(); }
    void copyTo(const class _OutputArray &dst) const;
    unsigned char *ptr(int r = 0) { return data + (size_t)r * step; }
    const unsigned char *ptr(int r = 0) const { return data + (size_t)r * step; }

private:
    int type_ = 0;
    std::shared_ptr<unsigned char> own_;
};

class _InputArray {
public:
    _InputArray() {}
    _InputArray(const Mat &m) : m_(&m) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }

protected:
    const Mat *m_ = nullptr;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat &m) : _InputArray(m), out_(&m) {}
    void create(int r, int c, int type) const { if (out_) out_->create(r, c, type); }
    void release() const { if (out_) out_->release(); }
    Mat getMat() const { return out_ ? *out_ : Mat(); }
    Mat *out_ = nullptr;
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

inline void Mat::copyTo(const _OutputArray &dst) const
{
    dst.create(rows, cols, type_);
    if (dst.out_)
        for (int r = 0; r < rows; ++r) std::memcpy(dst.out_->ptr(r), ptr(r), (size_t)cols * elemSize());
}

struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
class KeyPoint {
public:
    KeyPoint() {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

}  // namespace cv
