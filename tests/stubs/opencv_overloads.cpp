// Type-checks (and, with a device, runs) the reference-signature overloads of include/sindyn_classes.hpp against the stub
// OpenCV header: ORB_SLAM2::DynaDetect::DynaDetect(const cv::InputArray&, ...), DetectDynaArea(const cv::InputArray&, const
// cv::InputArray&, cv::OutputArray&, cv::OutputArray&, int) (DynaDetect.h:98-131) and ORBextractor::operator()(cv::InputArray,
// cv::InputArray, std::vector<cv::KeyPoint>&, cv::OutputArray) (ORBextractor.h:62-64) -- written like the reference driver
// (rgbd_tum_noros.cc:100-139) writes them.
#define SINDYN_WITH_OPENCV
#include "sindyn_classes.hpp"

#include <cstdio>

int main(int argc, char **)
{
    const int W = 640, H = 480;
    cv::Mat imRGB(H, W, CV_8UC3), imD(H, W, CV_16UC1), imDynaMask, imLabel;
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            for (int k = 0; k < 3; ++k) imRGB.ptr(r)[3 * c + k] = (unsigned char)((r * 7 + c * 13 + k * 31) & 255);
            ((unsigned short *)imD.ptr(r))[c] = (unsigned short)(5000 + ((r / 16 + c / 16) & 3) * 500);
        }
    if (argc > 1000) return 0;
    try {
        // rgbd_tum_noros.cc:106-107
        std::shared_ptr<ORB_SLAM2::DynaDetect> detertor = std::make_shared<ORB_SLAM2::DynaDetect>(imRGB, imRGB, 535.4f, 539.2f, 320.1f, 247.6f, 5000.0f);
        detertor->DetectDynaArea(imRGB, imD, imDynaMask, imLabel, 1);      // :135
        ORB_SLAM2::ORBextractor extractor(1000, 1.2f, 8, 20, 7);          // Tracking.cc:113-125
        std::vector<cv::KeyPoint> keys;
        cv::Mat gray(H, W, CV_8UC1), desc;
        for (int r = 0; r < H; ++r)
            for (int c = 0; c < W; ++c) gray.ptr(r)[c] = imRGB.ptr(r)[3 * c + 1];
        extractor(gray, imDynaMask, keys, desc);                          // Frame.cc:300-317
        std::printf("ok: mask %dx%d label %dx%d keypoints %zu descriptors %dx%d\n", imDynaMask.cols, imDynaMask.rows, imLabel.cols, imLabel.rows, keys.size(),
                    desc.cols, desc.rows);
        return (imDynaMask.rows == H && imLabel.cols == W && (int)keys.size() == desc.rows) ? 0 : 2;
    } catch (const sindyn::Error &e) {
        std::printf("sindyn::Error: %s\n", e.what());      // no device: the library fails loudly (there is no CPU fallback)
        return 3;
    }
}
