// Qt-free stand-in for the reference's ui/FrameBuffer.h (TEST INFRASTRUCTURE ONLY).
//
// core/Integrator.h:11 of the reference includes "ui/FrameBuffer.h", and the real
// header derives from QObject (ui/FrameBuffer.h:6-9); Qt is not installed here, so the
// oracle build puts this directory first on the include path.  Only the members that
// SamplerIntegrator::Render touches are provided (core/Integrator.cpp:230,307-310):
// renderCountIncrease, update_f_u_c (running mean over Render() calls + the exposure
// tonemap of ui/FrameBuffer.h:127-149) and set_uc, plus InitBuffer and read accessors
// for the harness.
#ifndef GNX_ORACLE_SHIM_FRAMEBUFFER_H
#define GNX_ORACLE_SHIM_FRAMEBUFFER_H

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

class FrameBuffer {
  public:
    FrameBuffer() {}
    void InitBuffer(int w = 800, int h = 600, int c = 4) {
        width_ = w; height_ = h; chan_ = c;
        u8_.assign((size_t)w * h * c, 0);
        f32_.assign((size_t)w * h * c, 0.f);
    }
    void renderCountIncrease() { ++passes_; }
    void renderCountClear() { passes_ = 0; }
    bool set_uc(int x, int y, int c, const unsigned char &v) {
        if (!inside(x, y)) return false;
        u8_[at(x, y, c)] = v;
        return true;
    }
    bool set_fc(int x, int y, int c, const float &v) {
        if (!inside(x, y)) return false;
        f32_[at(x, y, c)] = v;
        return true;
    }
    bool update_f_u_c(int x, int y, int c, const float &v) {
        if (!inside(x, y)) return false;
        size_t o = at(x, y, c);
        float w = 1.0f / (float)passes_;
        f32_[o] = w * v + (1.0f - w) * f32_[o];
        float exposure = 0.75;
        float mapped = 1.0f - expf(-f32_[o] * 1.0f / (1 - exposure));
        u8_[o] = mapped * 255;
        return true;
    }
    unsigned char *getUCbuffer() { return u8_.data(); }
    float *getFbuffer() { return f32_.data(); }
    int getWidth() const { return width_; }
    int getHeight() const { return height_; }
    void saveToFile(const std::string &) {}

  private:
    bool inside(int x, int y) const {
        return !f32_.empty() && x >= 0 && y >= 0 && x < width_ && y < height_;
    }
    size_t at(int x, int y, int c) const { return ((size_t)x + (size_t)y * width_) * chan_ + c; }
    std::vector<unsigned char> u8_;
    std::vector<float> f32_;
    int width_ = 0, height_ = 0, chan_ = 0, passes_ = 0;
};

#endif
