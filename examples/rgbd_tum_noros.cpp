// rgbd_tum_noros.cpp -- the reference's non-ROS RGB-D driver loop (ORB_SLAM2/Examples/RGB-D/rgbd_tum_noros.cc:37-215)
// reduced to the hot path this repository implements: LoadImages (:217-242), DynaDetect construction from frame 0
// (:103-107), per frame DetectDynaArea -> 15x15 ellipse dilation (:132-139) -> the masked ORB extraction that
// System::TrackRGBD -> Tracking::GrabImageRGBD -> Frame::ExtractORB2 performs (Tracking.cc:246-269, Frame.cc:300-309),
// and the timing summary (:198-209).  ORB-SLAM2 tracking / mapping themselves are out of scope.
//
//   rgbd_tum_noros <vocabulary (ignored)> <settings.yaml> <sequence dir> <associations.txt> [output dir]
//
// OpenCV is not available in the build image: images are read as binary PPM (rgb, 8-bit P6 in B,G,R byte order as
// cv::imread would deliver) and 16-bit PGM (depth, P5, big endian); sindslam_b200.synth writes both next to the PNGs.
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../include/sindyn_classes.hpp"

using namespace std;

static void LoadImages(const string &strAssociationFilename, vector<string> &rgb, vector<string> &depth, vector<double> &ts)
{
    ifstream f(strAssociationFilename.c_str());
    string s;
    while (getline(f, s)) {
        if (s.empty()) continue;
        stringstream ss(s);
        double t; string a, b;
        ss >> t >> a >> t >> b;
        ts.push_back(t); rgb.push_back(a); depth.push_back(b);
    }
}

static string swap_ext(string p, const char *ext) { size_t d = p.rfind('.'); return (d == string::npos ? p : p.substr(0, d)) + ext; }

static bool read_pnm(const string &path, vector<uint8_t> &buf, int &w, int &h, int &ch, int &bytes)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    char magic[3] = {0};
    int maxv = 0;
    if (fscanf(fp, "%2s %d %d %d", magic, &w, &h, &maxv) != 4) { fclose(fp); return false; }
    fgetc(fp);
    ch = magic[1] == '6' ? 3 : 1;
    bytes = maxv > 255 ? 2 : 1;
    buf.resize((size_t)w * h * ch * bytes);
    bool ok = fread(buf.data(), 1, buf.size(), fp) == buf.size();
    fclose(fp);
    if (ok && bytes == 2)
        for (size_t i = 0; i + 1 < buf.size(); i += 2) std::swap(buf[i], buf[i + 1]);   // big endian -> host order
    return ok;
}

static void write_pgm(const string &path, const sindyn::Image &img)
{
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) return;
    fprintf(fp, "P5\n%d %d\n255\n", img.cols, img.rows);
    fwrite(img.buf.data(), 1, img.buf.size(), fp);
    fclose(fp);
}

static float yaml_value(const string &path, const string &key, float def)
{
    ifstream f(path.c_str());
    string line;
    while (getline(f, line)) {
        size_t p = line.find(key + ":");
        if (p == 0) return (float)atof(line.substr(key.size() + 1).c_str());
    }
    return def;
}

int main(int argc, char **argv)
{
    if (argc < 5) {
        cerr << endl << "Usage: ./rgbd_tum_noros path_to_vocabulary path_to_settings path_to_sequence path_to_association [output_dir]" << endl;
        return 1;
    }
    const string seq = argv[3], outdir = argc > 5 ? argv[5] : "";
    vector<string> vRGB, vD;
    vector<double> vT;
    LoadImages(argv[4], vRGB, vD, vT);
    const int nImages = (int)vRGB.size();
    if (vRGB.empty()) { cerr << endl << "No images found in provided path." << endl; return 1; }
    // rgbd_tum_noros.cc:76-86 and Tracking.cc:113-119
    const float fx = yaml_value(argv[2], "Camera.fx", 535.4f), fy = yaml_value(argv[2], "Camera.fy", 539.2f);
    const float cx = yaml_value(argv[2], "Camera.cx", 320.1f), cy = yaml_value(argv[2], "Camera.cy", 247.6f);
    const float depthScale = yaml_value(argv[2], "DepthMapFactor", 5000.0f);
    const int rgbOrder = (int)yaml_value(argv[2], "Camera.RGB", 1.0f);
    const int nFeatures = (int)yaml_value(argv[2], "ORBextractor.nFeatures", 1500), nLevels = (int)yaml_value(argv[2], "ORBextractor.nLevels", 8);
    const float fScale = yaml_value(argv[2], "ORBextractor.scaleFactor", 1.2f);
    const int iniTh = (int)yaml_value(argv[2], "ORBextractor.iniThFAST", 15), minTh = (int)yaml_value(argv[2], "ORBextractor.minThFAST", 5);

    vector<uint8_t> rgb, dep;
    int w, h, ch, bytes;
    if (!read_pnm(seq + "/" + swap_ext(vRGB[0], ".ppm"), rgb, w, h, ch, bytes) || ch != 3) { cerr << "Failed to load image at: " << vRGB[0] << endl; return 1; }
    sindyn::ImageView first(rgb.data(), h, w, (size_t)w * 3, 3, 1);
    try {
        std::shared_ptr<ORB_SLAM2::DynaDetect> detertor = std::make_shared<ORB_SLAM2::DynaDetect>(first, first, fx, fy, cx, cy, depthScale);
        ORB_SLAM2::ORBextractor extractor(nFeatures, fScale, nLevels, iniTh, minTh);
        extractor.keepPyramidOnHost = false;
        sindyn::Image imDynaMask, imLabel, desc;
        imDynaMask.create(h, w); imLabel.create(h, w);   // all-zero mask for frame 0 (rgbd_tum_noros.cc:100-101)
        vector<uint8_t> gray((size_t)w * h);
        vector<sindyn::KeyPoint> kps;
        vector<float> vTimesDynamic(nImages, 0.f), vTimesOrb(nImages, 0.f);
        for (int ni = 0; ni < nImages; ni++) {
            cout << "----------now processing " << ni << " img-----------------" << endl;
            int dw, dh, dch, dbytes;
            if (!read_pnm(seq + "/" + swap_ext(vRGB[ni], ".ppm"), rgb, w, h, ch, bytes) || !read_pnm(seq + "/" + swap_ext(vD[ni], ".pgm"), dep, dw, dh, dch, dbytes) ||
                dbytes != 2) { cerr << endl << "Failed to load image at: " << seq << "/" << vRGB[ni] << endl; return 1; }
            sindyn::ImageView imRGB(rgb.data(), h, w, (size_t)w * 3, 3, 1), imD(dep.data(), dh, dw, (size_t)dw * 2, 1, 2);
            auto t0 = chrono::steady_clock::now();
            if (ni >= 1) {
                detertor->DetectDynaArea(imRGB, imD, imDynaMask, imLabel, ni);
                if (!imDynaMask.empty()) detertor->morphologyExEllipse(imDynaMask, 15, 0);
            }
            auto t1 = chrono::steady_clock::now();
            // Tracking::GrabImageRGBD: cvtColor(mImGray, RGB2GRAY | BGR2GRAY) (Tracking.cc:246-258), 15-bit fixed point
            for (size_t i = 0; i < gray.size(); ++i) {
                const int c0 = rgb[3 * i], c1 = rgb[3 * i + 1], c2 = rgb[3 * i + 2];
                gray[i] = (uint8_t)(rgbOrder ? ((c0 * 9798 + c1 * 19235 + c2 * 3735 + 16384) >> 15) : ((c0 * 3735 + c1 * 19235 + c2 * 9798 + 16384) >> 15));
            }
            extractor(sindyn::ImageView(gray.data(), h, w, (size_t)w, 1, 1), imDynaMask, kps, desc);
            auto t2 = chrono::steady_clock::now();
            vTimesDynamic[ni] = chrono::duration_cast<chrono::duration<float>>(t1 - t0).count();
            vTimesOrb[ni] = chrono::duration_cast<chrono::duration<float>>(t2 - t1).count();
            size_t dyn = 0;
            for (uint8_t v : imDynaMask.buf) dyn += v == 255;
            cout << "dynamic px " << dyn << ", keypoints " << kps.size() << endl;
            if (!outdir.empty()) {
                char name[64];
                snprintf(name, sizeof name, "/%06d", ni);
                write_pgm(outdir + name + "_mask.pgm", imDynaMask);
                write_pgm(outdir + name + "_label.pgm", imLabel);
            }
        }
        float tot = 0, tot2 = 0;
        for (int ni = 0; ni < nImages; ni++) { tot += vTimesDynamic[ni]; tot2 += vTimesOrb[ni]; }
        cout << "-------" << endl << endl;
        cout << "mean dynamic detecting time: " << tot / nImages << endl;
        cout << "mean ORB extraction time: " << tot2 / nImages << endl;
    } catch (const sindyn::Error &e) {
        cerr << "sindyn error " << e.status << ": " << e.what() << endl;
        return 2;
    }
    return 0;
}
