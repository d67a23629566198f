// Minimal stand-in for <opencv2/core.hpp>: only what the class layer / adapters use.  See ../README.md.
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr); } while (0)

namespace cv {

struct Point2f { float x = 0, y = 0; Point2f() = default; Point2f(float x_, float y_) : x(x_), y(y_) {} };

struct KeyPoint {
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
};

class _OutputArray;

class Mat {
public:
    int rows = 0, cols = 0;
    std::size_t step = 0;
    unsigned char* data = nullptr;
    Mat() = default;
    Mat(int rows_, int cols_, int /*type: CV_8U only*/) : rows(rows_), cols(cols_), step((std::size_t)cols_) {
        store_ = std::make_shared<std::vector<unsigned char>>((std::size_t)rows_ * cols_, 0);
        data = store_->data();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    Mat rowRange(int a, int b) const { Mat m = *this; m.rows = b - a; m.data = data + (std::size_t)a * step; return m; }
    Mat row(int i) const { return rowRange(i, i + 1); }
    Mat clone() const {
        Mat m(rows, cols, CV_8U);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (std::size_t)r * m.step, data + (std::size_t)r * step, (std::size_t)cols);
        return m;
    }
    inline void copyTo(const _OutputArray& dst) const;
    unsigned char* ptr(int r) { return data + (std::size_t)r * step; }
    const unsigned char* ptr(int r) const { return data + (std::size_t)r * step; }
private:
    std::shared_ptr<std::vector<unsigned char>> store_;
};

class _InputArray {
public:
    _InputArray() = default;
    _InputArray(const Mat& m) : m_(&m) {}   // NOLINT: implicit, as in OpenCV
    bool empty() const { return m_ == nullptr || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
private:
    const Mat* m_ = nullptr;
};

class _OutputArray {
public:
    _OutputArray(Mat& m) : m_(&m) {}        // NOLINT
    void release() const { *m_ = Mat(); }
    void assign(const Mat& src) const { *m_ = src.clone(); }
private:
    Mat* m_;
};

inline void Mat::copyTo(const _OutputArray& dst) const { dst.assign(*this); }
inline _InputArray noArray() { return _InputArray(); }
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

}  // namespace cv
