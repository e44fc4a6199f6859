// Minimal row-major fp32 matrix standing in for the reference's
//   typedef Eigen::Matrix<float, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> Mat;   (env/env.hpp:14)
// Only what the reference's call sites use on the hot path's host side (ppo2.cpp:44-64, runner.hpp, env/*.hpp):
// Mat(r,c), (i,j), rows()/cols()/data(), Zero/Ones, row access, operator<<.  No Eigen anywhere.
#ifndef PPO_B200_MAT_HPP
#define PPO_B200_MAT_HPP

#include <cassert>
#include <cstring>
#include <iostream>
#include <vector>

class Mat {
public:
    Mat() : r_(0), c_(0) {}
    Mat(int rows, int cols) : r_(rows), c_(cols), v_(static_cast<size_t>(rows) * cols) {}
    static Mat Zero(int rows, int cols) {
        Mat m(rows, cols);
        std::fill(m.v_.begin(), m.v_.end(), 0.f);
        return m;
    }
    static Mat Ones(int rows, int cols) { return Constant(rows, cols, 1.f); }
    static Mat Constant(int rows, int cols, float value) {
        Mat m(rows, cols);
        std::fill(m.v_.begin(), m.v_.end(), value);
        return m;
    }
    int rows() const { return r_; }
    int cols() const { return c_; }
    size_t size() const { return v_.size(); }
    float* data() { return v_.data(); }
    const float* data() const { return v_.data(); }
    float& operator()(int i, int j) {
        assert(i >= 0 && i < r_ && j >= 0 && j < c_);
        return v_[static_cast<size_t>(i) * c_ + j];
    }
    float operator()(int i, int j) const {
        assert(i >= 0 && i < r_ && j >= 0 && j < c_);
        return v_[static_cast<size_t>(i) * c_ + j];
    }
    Mat row(int i) const {
        Mat m(1, c_);
        std::memcpy(m.data(), v_.data() + static_cast<size_t>(i) * c_, sizeof(float) * c_);
        return m;
    }
    void set_row(int i, const Mat& src) {
        assert(src.size() == static_cast<size_t>(c_));
        std::memcpy(v_.data() + static_cast<size_t>(i) * c_, src.data(), sizeof(float) * c_);
    }
    Mat operator*(float s) const {
        Mat m(*this);
        for (auto& x : m.v_) x *= s;
        return m;
    }
    float sum() const {
        float s = 0.f;
        for (float x : v_) s += x;
        return s;
    }
    float squaredNorm() const {
        float s = 0.f;
        for (float x : v_) s += x * x;
        return s;
    }
    Mat operator-(const Mat& o) const {
        assert(r_ == o.r_ && c_ == o.c_);
        Mat m(*this);
        for (size_t i = 0; i < v_.size(); ++i) m.v_[i] -= o.v_[i];
        return m;
    }
    Mat col(int j) const {
        Mat m(r_, 1);
        for (int i = 0; i < r_; ++i) m(i, 0) = (*this)(i, j);
        return m;
    }

private:
    int r_, c_;
    std::vector<float> v_;
};

inline Mat operator*(float s, const Mat& m) { return m * s; }

inline std::ostream& operator<<(std::ostream& os, const Mat& m) {
    for (int i = 0; i < m.rows(); ++i) {
        for (int j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}

#endif
