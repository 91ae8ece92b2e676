// sindyn_classes.hpp -- header-only C++ mirror of the two reference classes that sit on the hot path,
// implemented on top of the C ABI of libsindyn_cuda (include/sindyn.h):
//
//   ORB_SLAM2::DynaDetect    <- ORB_SLAM2/include/DynaDetect.h:95-189   (ctor :98-126, DetectDynaArea :127-131)
//   ORB_SLAM2::ORBextractor  <- ORB_SLAM2/include/ORBextractor.h:47-114 (ctor :54-55, operator() :62-64, getters :66-86,
//                                                                        public mvImagePyramid :88)
//
// Same class names, method names, argument order and meaning, and error behaviour (void returns; failure = empty
// output images or a thrown exception -- the reference throws cv::Exception from CV_Assert; operator() returns
// silently on an empty image, ORBextractor.cc:1046-1047).
//
// The reference's signatures take cv::InputArray / cv::OutputArray.  OpenCV's C++ headers do not exist in the build
// image, so the classes are written against a minimal image view (sindyn::ImageView = {data, rows, cols, step,
// channels, elem size}) and an owning output image (sindyn::Image).  Define SINDYN_WITH_OPENCV before including this
// header to get the exact cv::InputArray / cv::OutputArray / cv::KeyPoint overloads as well.
#ifndef SINDYN_CLASSES_HPP
#define SINDYN_CLASSES_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "sindyn.h"

#ifdef SINDYN_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace sindyn {

struct ImageView {   // what cv::InputArray::getMat() gives the reference: a borrowed, possibly padded 2-D array
    const void *data = nullptr;
    int rows = 0, cols = 0;
    size_t step = 0;       // bytes per row (cv::Mat::step)
    int channels = 1;
    int elem_bytes = 1;    // bytes per channel (1 = 8U, 2 = 16U)
    ImageView() {}
    ImageView(const void *d, int r, int c, size_t s, int ch, int eb) : data(d), rows(r), cols(c), step(s), channels(ch), elem_bytes(eb) {}
    bool empty() const { return !data || rows <= 0 || cols <= 0; }
#ifdef SINDYN_WITH_OPENCV
    explicit ImageView(const cv::Mat &m) : data(m.data), rows(m.rows), cols(m.cols), step(m.step), channels(m.channels()), elem_bytes((int)m.elemSize1()) {}
#endif
};

struct Image {   // owning 8U output (the cv::OutputArray role); dense rows
    std::vector<uint8_t> buf;
    int rows = 0, cols = 0, channels = 1;
    void create(int r, int c, int ch = 1) { rows = r; cols = c; channels = ch; buf.assign((size_t)r * c * ch, 0); }
    void release() { buf.clear(); rows = cols = 0; }
    bool empty() const { return buf.empty(); }
    size_t step() const { return (size_t)cols * channels; }
    uint8_t *ptr(int r = 0) { return buf.data() + (size_t)r * step(); }
    const uint8_t *ptr(int r = 0) const { return buf.data() + (size_t)r * step(); }
    operator ImageView() const { return ImageView(buf.data(), rows, cols, step(), channels, 1); }
#ifdef SINDYN_WITH_OPENCV
    cv::Mat mat() { return cv::Mat(rows, cols, CV_8UC(channels), buf.data()); }
#endif
};

struct Point2f { float x = 0, y = 0; };
struct KeyPoint {   // the cv::KeyPoint fields ORB-SLAM2 reads
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

class Error : public std::runtime_error {
public:
    int status;
    Error(int st, const std::string &what) : std::runtime_error(what), status(st) {}
};

}  // namespace sindyn

namespace ORB_SLAM2 {

class DynaDetect {
public:
    // DynaDetect.h:98-105.  imgLast_ / imgLastLast_: 8UC3 BGR frames (the driver passes frame 0 twice,
    // rgbd_tum_noros.cc:103-107).  The image size fixes the handle's geometry (the reference hard-codes 640x480).
    DynaDetect(const sindyn::ImageView &imgLast_, const sindyn::ImageView &imgLastLast_, float fx_, float fy_, float cx_, float cy_,
               float depthScale_, const sindyn_config *overrides = nullptr)
    {
        if (imgLast_.empty() || imgLastLast_.empty() || imgLast_.channels != 3 || imgLast_.elem_bytes != 1 || imgLast_.rows != imgLastLast_.rows ||
            imgLast_.cols != imgLastLast_.cols)
            throw sindyn::Error(SINDYN_ERR_INVALID, "DynaDetect: imgLast / imgLastLast must be non-empty 8UC3 images of one size");
        sindyn_config cfg;
        if (overrides) cfg = *overrides;
        else sindyn_default_config(&cfg, imgLast_.cols, imgLast_.rows);
        cfg.width = imgLast_.cols; cfg.height = imgLast_.rows;
        cfg.fx = fx_; cfg.fy = fy_; cfg.cx = cx_; cfg.cy = cy_; cfg.depth_scale = depthScale_;
        check(sindyn_create(&cfg, &h_), "sindyn_create");
        check(sindyn_set_prev_frames(h_, (const uint8_t *)imgLast_.data, imgLast_.step, (const uint8_t *)imgLastLast_.data, imgLastLast_.step),
              "sindyn_set_prev_frames");
        width_ = cfg.width; height_ = cfg.height;
    }
    ~DynaDetect() { if (h_) sindyn_destroy(h_); }
    DynaDetect(const DynaDetect &) = delete;
    DynaDetect &operator=(const DynaDetect &) = delete;

    // DynaDetect.h:127-131 / DynaDetect.cc:1377-1666.  img_: 8UC3 BGR; imgDepth_: 16UC1 raw depth;
    // imgDyna_: 8UC1 {0 invalid depth, 125 static, 255 dynamic}; imgLabel_: 8UC1 merged cluster ids; nImg_: frame index.
    void DetectDynaArea(const sindyn::ImageView &img_, const sindyn::ImageView &imgDepth_, sindyn::Image &imgDyna_, sindyn::Image &imgLabel_,
                        int nImg_)
    {
        if (img_.empty() || imgDepth_.empty()) { imgDyna_.release(); imgLabel_.release(); return; }   // the driver checks .empty()
        if (img_.rows != height_ || img_.cols != width_ || img_.channels != 3 || imgDepth_.rows != height_ || imgDepth_.cols != width_ ||
            imgDepth_.elem_bytes != 2)
            throw sindyn::Error(SINDYN_ERR_INVALID, "DetectDynaArea: img must be 8UC3 and imgDepth 16UC1 of the constructor's size");
        imgDyna_.create(height_, width_);
        imgLabel_.create(height_, width_);
        check(sindyn_detect(h_, (const uint8_t *)img_.data, img_.step, (const uint16_t *)imgDepth_.data, imgDepth_.step, imgDyna_.ptr(),
                            imgDyna_.step(), imgLabel_.ptr(), imgLabel_.step(), nImg_),
              "sindyn_detect");
    }

#ifdef SINDYN_WITH_OPENCV
    DynaDetect(const cv::InputArray &imgLast_, const cv::InputArray &imgLastLast_, float fx_, float fy_, float cx_, float cy_, float depthScale_)
        : DynaDetect(sindyn::ImageView(imgLast_.getMat()), sindyn::ImageView(imgLastLast_.getMat()), fx_, fy_, cx_, cy_, depthScale_) {}
    void DetectDynaArea(const cv::InputArray &img_, const cv::InputArray &imgDepth_, cv::OutputArray &imgDyna_, cv::OutputArray &imgLabel_, int nImg_)
    {
        sindyn::Image dyna, label;
        DetectDynaArea(sindyn::ImageView(img_.getMat()), sindyn::ImageView(imgDepth_.getMat()), dyna, label, nImg_);
        if (dyna.empty()) { imgDyna_.release(); imgLabel_.release(); return; }
        dyna.mat().copyTo(imgDyna_);     // deep copies, like DynaDetect.cc:1635-1636
        label.mat().copyTo(imgLabel_);
    }
#endif

    // cv::morphologyEx(img, img, op, getStructuringElement(MORPH_ELLIPSE, k x k)) on the device -- the driver's
    // post-step (rgbd_tum_noros.cc:108,136-139) without an OpenCV dependency. op: 0 dilate, 1 erode, 2 open, 3 close.
    void morphologyExEllipse(sindyn::Image &img, int k, int op = 0)
    {
        if (img.empty()) return;
        std::vector<uint8_t> out(img.buf.size());
        check(sindyn_morph_ellipse(h_, img.ptr(), img.step(), out.data(), img.step(), img.cols, img.rows, k, op), "sindyn_morph_ellipse");
        img.buf.swap(out);
    }

    // per-stage device milliseconds of the last call (the reference prints running means, DynaDetect.cc:1644-1649)
    std::vector<float> stageMilliseconds() const { std::vector<float> ms(16, 0.f); sindyn_get_stage_ms(h_, ms.data(), 16); return ms; }
    sindyn_handle handle() const { return h_; }

private:
    void check(int st, const char *what) const
    {
        if (st != SINDYN_OK) throw sindyn::Error(st, std::string(what) + ": " + (h_ ? sindyn_last_error(h_) : "no handle (is a CUDA device present?)"));
    }
    sindyn_handle h_ = nullptr;
    int width_ = 0, height_ = 0;
};

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };   // ORBextractor.h:52 (HARRIS_SCORE is unused by the reference too)

    // ORBextractor.h:54-55 / ORBextractor.cc:410-470 -- scale tables are host-side so the getters work before the first frame.
    ORBextractor(int nfeatures_, float scaleFactor_, int nlevels_, int iniThFAST_, int minThFAST_, int device = 0)
        : nfeatures(nfeatures_), scaleFactor(scaleFactor_), nlevels(nlevels_), iniThFAST(iniThFAST_), minThFAST(minThFAST_), device_(device)
    {
        mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
        mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
        for (int i = 1; i < nlevels; i++) {
            mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);
            mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
        }
        mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        for (int i = 0; i < nlevels; i++) {
            mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
            mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
        }
        mvImagePyramid.resize(nlevels);
    }
    ~ORBextractor() { if (h_) sindyn_orb_destroy(h_); }
    ORBextractor(const ORBextractor &) = delete;
    ORBextractor &operator=(const ORBextractor &) = delete;

    // ORBextractor.h:62-64 / ORBextractor.cc:1043-1164.  image: 8UC1; mask: 8UC1 dynamic mask or empty; keypoints with
    // mask == 255 at their (scaled) position are erased unless fewer than 250 would remain.
    void operator()(const sindyn::ImageView &image, const sindyn::ImageView &mask, std::vector<sindyn::KeyPoint> &keypoints,
                    sindyn::Image &descriptors)
    {
        if (image.empty()) return;
        if (image.channels != 1 || image.elem_bytes != 1) throw sindyn::Error(SINDYN_ERR_INVALID, "ORBextractor: image must be 8UC1");   // assert(image.type() == CV_8UC1)
        ensure(image.cols, image.rows);
        const int cap = nfeatures * 2 + 64;
        kp_.resize(cap);
        desc_.resize((size_t)cap * 32);
        int n = 0;
        const bool have_mask = !mask.empty();
        if (have_mask && (mask.rows != image.rows || mask.cols != image.cols)) throw sindyn::Error(SINDYN_ERR_INVALID, "ORBextractor: mask size");
        int st = sindyn_orb_extract(h_, (const uint8_t *)image.data, image.step, have_mask ? (const uint8_t *)mask.data : nullptr, have_mask ? mask.step : 0,
                                    kp_.data(), desc_.data(), cap, &n);
        if (st != SINDYN_OK) throw sindyn::Error(st, std::string("sindyn_orb_extract: ") + sindyn_orb_last_error(h_));
        keypoints.clear();
        keypoints.reserve(n);
        for (int i = 0; i < n; ++i) {
            sindyn::KeyPoint k;
            k.pt.x = kp_[i].x; k.pt.y = kp_[i].y; k.size = kp_[i].size; k.angle = kp_[i].angle; k.response = kp_[i].response; k.octave = kp_[i].octave;
            keypoints.push_back(k);
        }
        if (n == 0) descriptors.release();
        else { descriptors.create(n, 32); std::memcpy(descriptors.ptr(), desc_.data(), (size_t)n * 32); }
        if (keepPyramidOnHost) {   // upstream ORB-SLAM2 code may read the public member
            for (int l = 0; l < nlevels; ++l) {
                int w = 0, hh = 0;
                sindyn_orb_get_pyramid_level(h_, l, nullptr, &w, &hh);
                mvImagePyramid[l].create(hh, w);
                sindyn_orb_get_pyramid_level(h_, l, mvImagePyramid[l].ptr(), &w, &hh);
            }
        }
    }

#ifdef SINDYN_WITH_OPENCV
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints, cv::OutputArray descriptors)
    {
        if (image.empty()) return;
        std::vector<sindyn::KeyPoint> k;
        sindyn::Image d;
        cv::Mat m = mask.empty() ? cv::Mat() : mask.getMat();
        (*this)(sindyn::ImageView(image.getMat()), m.empty() ? sindyn::ImageView() : sindyn::ImageView(m), k, d);
        keypoints.clear();
        for (const auto &p : k) keypoints.emplace_back(p.pt.x, p.pt.y, p.size, p.angle, p.response, p.octave, -1);
        if (d.empty()) descriptors.release();
        else d.mat().copyTo(descriptors);
    }
#endif

    // what Frame::Frame does with the keypoints right after the extraction (Frame.cc:143-170), computed on
    // the device while the keypoints are still resident: mvKeysUn, mvDepth, mvuRight, image bounds and the 64 x 48 mGrid (CSR).
    struct FrameFeatures {
        std::vector<float> keysUn, depth, uRight;   // n x 2, n, n
        float bounds[4];                            // mnMinX, mnMaxX, mnMinY, mnMaxY
        std::vector<int> gridOffsets, gridIndices;  // mGrid[i][j] = gridIndices[gridOffsets[i*48+j] .. gridOffsets[i*48+j+1])
    };
    void ComputeFrameFeatures(const sindyn::ImageView &imDepthRaw, const sindyn_frame_params &p, FrameFeatures &out)
    {
        if (!h_) throw sindyn::Error(SINDYN_ERR_STATE, "ComputeFrameFeatures: call operator() first");
        const int cap = nfeatures * 2 + 64;
        out.keysUn.resize((size_t)cap * 2); out.depth.resize(cap); out.uRight.resize(cap);
        out.gridOffsets.resize(64 * 48 + 1); out.gridIndices.resize(cap);
        int n = 0;
        int st = sindyn_orb_frame_features(h_, (const uint16_t *)imDepthRaw.data, imDepthRaw.step, &p, out.keysUn.data(), out.depth.data(),
                                           out.uRight.data(), out.bounds, out.gridOffsets.data(), out.gridIndices.data(), cap, &n);
        if (st != SINDYN_OK) throw sindyn::Error(st, std::string("sindyn_orb_frame_features: ") + sindyn_orb_last_error(h_));
        out.keysUn.resize((size_t)n * 2); out.depth.resize(n); out.uRight.resize(n); out.gridIndices.resize(out.gridOffsets.back());
    }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }
    // the C-ABI handle (the frame of the last operator() / ComputeFrameFeatures call is resident in it) and its key point capacity
    sindyn_orb_handle handle() const { return h_; }
    int capacity() const { return nfeatures * 2 + 64; }

    std::vector<sindyn::Image> mvImagePyramid;   // ORBextractor.h:88
    bool keepPyramidOnHost = true;               // set to false to skip the device-to-host copy of the 8 levels

protected:
    void ensure(int w, int hgt)
    {
        if (h_ && w == w_ && hgt == h_img_) return;
        if (h_) { sindyn_orb_destroy(h_); h_ = nullptr; }
        int st = sindyn_orb_create(nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, w, hgt, device_, &h_);
        if (st != SINDYN_OK) {
            std::string m = h_ ? sindyn_orb_last_error(h_) : "no handle (is a CUDA device present?)";
            if (h_) { sindyn_orb_destroy(h_); h_ = nullptr; }
            throw sindyn::Error(st, "sindyn_orb_create: " + m);
        }
        w_ = w; h_img_ = hgt;
    }
    int nfeatures;
    double scaleFactor;   // ORBextractor.h:100: a double holding the float argument
    int nlevels, iniThFAST, minThFAST;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    sindyn_orb_handle h_ = nullptr;
    int device_ = 0, w_ = 0, h_img_ = 0;
    std::vector<sindyn_keypoint> kp_;
    std::vector<uint8_t> desc_;
};

// Mirrors ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
// (include/ORBmatcher.h:52-54, src/ORBmatcher.cc:1328-1470).  CurrentFrame is the frame resident in `extractor` (operator() +
// ComputeFrameFeatures); LastFrame's map points are handed over as arrays (see sindyn_orb_search_by_projection).
class ORBmatcher {
public:
    static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;   // ORBmatcher.cc:37-39
    ORBmatcher(float nnratio = 0.6f, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    struct LastFramePoints {   // one entry per key point of LastFrame
        std::vector<float> xyz_world;        // 3 per point: pMP->GetWorldPos()
        std::vector<uint8_t> valid;          // mvpMapPoints[i] != NULL && !mvbOutlier[i]
        std::vector<uint8_t> desc;           // 32 per point: pMP->GetDescriptor()
        std::vector<int> octave;             // mvKeys[i].octave
        std::vector<float> angle;            // mvKeysUn[i].angle
        std::vector<uint8_t> observed;       // pMP->Observations() > 0
    };
    // params: intrinsics / mbf / mb of CurrentFrame and both poses (mono and check_orientation are filled in here).
    // vMatches[i2] = index into LastFrame of the map point assigned to current key point i2, or -1.  Returns nmatches.
    int SearchByProjection(ORBextractor &extractor, const LastFramePoints &LastFrame, sindyn_match_params params, const float th, const bool bMono,
                           std::vector<int> &vMatches, const std::vector<uint8_t> *currentBlocked = nullptr)
    {
        const int n_last = (int)LastFrame.valid.size();
        if ((int)LastFrame.xyz_world.size() != 3 * n_last || (int)LastFrame.desc.size() != 32 * n_last || (int)LastFrame.octave.size() != n_last ||
            (int)LastFrame.angle.size() != n_last || (int)LastFrame.observed.size() != n_last)
            throw sindyn::Error(SINDYN_ERR_INVALID, "SearchByProjection: LastFrame arrays disagree in length");
        params.th = th; params.mono = bMono ? 1 : 0; params.check_orientation = mbCheckOrientation ? 1 : 0;
        vMatches.assign((size_t)extractor.capacity(), -1);
        int n_cur = 0, nmatches = 0;
        const int st = sindyn_orb_search_by_projection(extractor.handle(), &params, n_last, LastFrame.xyz_world.data(), LastFrame.valid.data(),
                                                       LastFrame.desc.data(), LastFrame.octave.data(), LastFrame.angle.data(), LastFrame.observed.data(),
                                                       currentBlocked ? currentBlocked->data() : nullptr, vMatches.data(), (int)vMatches.size(), &n_cur,
                                                       &nmatches);
        if (st != SINDYN_OK) throw sindyn::Error(st, std::string("sindyn_orb_search_by_projection: ") + sindyn_orb_last_error(extractor.handle()));
        vMatches.resize((size_t)n_cur);
        return nmatches;
    }

protected:
    float mfNNratio;
    bool mbCheckOrientation;
};

}  // namespace ORB_SLAM2

namespace sindyn {

// Mirrors the two SubscribeAndPublish::generatePointCloud overloads of the dense-map consumer
// (octomap_pub/src/pubPointCloud.cc:392-470 and :471-678): same argument order and meaning; the result is returned instead of
// being left in the node's tempCloudOneFrame member, and imgDynaMaskNew (a local of the reference) is handed out on request.
// Poses are row-major 4 x 4 doubles (Eigen::Isometry3d::matrix() / Eigen::Matrix4d are column-major: pass the transpose or use
// the Eigen overload a binding adds).
class PointCloudGenerator {
public:
    PointCloudGenerator(int width, int height, double fx, double fy, double cx, double cy, double depthScale, int device = 0)
    {
        sindyn_config cfg;
        sindyn_default_config(&cfg, width, height);
        cfg.fx = (float)fx; cfg.fy = (float)fy; cfg.cx = (float)cx; cfg.cy = (float)cy; cfg.depth_scale = (float)depthScale;
        cfg.device = device;
        cfg.plane_edges = 0; cfg.refine = 0;
        intr_[0] = fx; intr_[1] = fy; intr_[2] = cx; intr_[3] = cy; intr_[4] = depthScale;
        int st = sindyn_create(&cfg, &h_);
        if (st != SINDYN_OK) throw Error(st, std::string("sindyn_create: ") + (h_ ? sindyn_last_error(h_) : "no handle (is a CUDA device present?)"));
        width_ = width; height_ = height;
    }
    ~PointCloudGenerator() { if (h_) sindyn_destroy(h_); }
    PointCloudGenerator(const PointCloudGenerator &) = delete;
    PointCloudGenerator &operator=(const PointCloudGenerator &) = delete;

    // pubPointCloud.cc:392-470
    std::vector<sindyn_point> generatePointCloud(const ImageView &imgRGB, const ImageView &imgDepth, const ImageView &imgDynaMask, const double *Twc)
    {
        need(imgRGB, 3, 1, "imgRGB"); need(imgDepth, 1, 2, "imgDepth"); need(imgDynaMask, 1, 1, "imgDynaMask");
        std::vector<sindyn_point> out((size_t)((height_ + 2) / 3) * ((width_ + 2) / 3));
        int n = 0;
        check(sindyn_cloud_single(h_, (const uint8_t *)imgRGB.data, imgRGB.step, (const uint16_t *)imgDepth.data, imgDepth.step,
                                  (const uint8_t *)imgDynaMask.data, imgDynaMask.step, Twc, intr_, out.data(), &n), "sindyn_cloud_single");
        out.resize((size_t)n);
        return out;
    }
    // pubPointCloud.cc:471-678
    std::vector<sindyn_point> generatePointCloud(const ImageView &imgRGB, const ImageView &imgDepth, const ImageView &imgDepthLast,
                                                 const ImageView &imgDynaMask, const ImageView &imgDynaMaskLast, const ImageView &imgLabel,
                                                 const double *poseRelative, const double *Twc, Image *imgDynaMaskNew = nullptr)
    {
        need(imgRGB, 3, 1, "imgRGB"); need(imgDepth, 1, 2, "imgDepth"); need(imgDepthLast, 1, 2, "imgDepthLast");
        need(imgDynaMask, 1, 1, "imgDynaMask"); need(imgDynaMaskLast, 1, 1, "imgDynaMaskLast"); need(imgLabel, 1, 1, "imgLabel");
        std::vector<sindyn_point> out((size_t)((height_ + 1) / 2) * ((width_ + 1) / 2));
        if (imgDynaMaskNew) imgDynaMaskNew->create(height_, width_);
        int n = 0;
        check(sindyn_cloud_consistent(h_, (const uint8_t *)imgRGB.data, imgRGB.step, (const uint16_t *)imgDepth.data, imgDepth.step,
                                      (const uint16_t *)imgDepthLast.data, imgDepthLast.step, (const uint8_t *)imgDynaMask.data, imgDynaMask.step,
                                      (const uint8_t *)imgDynaMaskLast.data, imgDynaMaskLast.step, (const uint8_t *)imgLabel.data, imgLabel.step,
                                      poseRelative, Twc, intr_, out.data(), &n, imgDynaMaskNew ? imgDynaMaskNew->ptr() : nullptr,
                                      imgDynaMaskNew ? imgDynaMaskNew->step() : 0, nullptr, nullptr),
              "sindyn_cloud_consistent");
        out.resize((size_t)n);
        return out;
    }

private:
    void need(const ImageView &v, int ch, int eb, const char *name) const
    {
        if (v.empty() || v.rows != height_ || v.cols != width_ || v.channels != ch || v.elem_bytes != eb)
            throw Error(SINDYN_ERR_INVALID, std::string("generatePointCloud: ") + name + " has the wrong size or type");
    }
    void check(int st, const char *what) const
    {
        if (st != SINDYN_OK) throw Error(st, std::string(what) + ": " + sindyn_last_error(h_));
    }
    sindyn_handle h_ = nullptr;
    int width_ = 0, height_ = 0;
    double intr_[5];
};

}  // namespace sindyn

#endif  // SINDYN_CLASSES_HPP
