// Qt-free stand-in for the reference's ui/FrameBuffer.h (TEST INFRASTRUCTURE ONLY).
//
// core/Integrator.h:11 of the reference includes "ui/FrameBuffer.h", and the real header derives from QObject
// (ui/FrameBuffer.h:6-9); Qt is not installed here, so the oracle build puts this directory first on the include path.
// The stand-in keeps the real class's DATA MEMBERS under their own names and in their own order (ubuffer, fbuffer,
// width, height, channals, curRenderCount — ui/FrameBuffer.h:153-160), because the product's bridge writes the two
// buffers directly (its translation unit sees the private members) and must compile unchanged against either header.
// Of the methods, the ones SamplerIntegrator::Render touches (core/Integrator.cpp:230,307-310) plus InitBuffer and read
// accessors for the harness.
#ifndef GNX_ORACLE_SHIM_FRAMEBUFFER_H
#define GNX_ORACLE_SHIM_FRAMEBUFFER_H

#include <cmath>
#include <cstring>
#include <string>

class FrameBuffer {
  public:
    FrameBuffer() : ubuffer(nullptr), fbuffer(nullptr), width(0), height(0), channals(0), curRenderCount(0) {}
    FrameBuffer(const FrameBuffer &) = delete;
    FrameBuffer &operator=(const FrameBuffer &) = delete;
    ~FrameBuffer() { Release(); }
    void InitBuffer(int w = 800, int h = 600, int c = 4) {
        Release();
        width = w; height = h; channals = c;
        const size_t n = (size_t)w * h * c;
        ubuffer = new unsigned char[n];
        fbuffer = new float[n];
        memset(ubuffer, 0, n);
        memset(fbuffer, 0, n * sizeof(float));
    }
    void renderCountIncrease() { curRenderCount++; }
    void renderCountClear() { curRenderCount = 0; }
    bool set_uc(int x, int y, int c, const unsigned char &v) {
        if (!inside(x, y)) return false;
        ubuffer[at(x, y, c)] = v;
        return true;
    }
    bool set_fc(int x, int y, int c, const float &v) {
        if (!inside(x, y)) return false;
        fbuffer[at(x, y, c)] = v;
        return true;
    }
    // running mean over Render() calls, then the exposure tonemap (ui/FrameBuffer.h:127-149)
    bool update_f_u_c(int x, int y, int c, const float &v) {
        if (!inside(x, y)) return false;
        size_t o = at(x, y, c);
        float w = 1.0f / (float)curRenderCount;
        fbuffer[o] = w * v + (1.0f - w) * fbuffer[o];
        float exposure = 0.75;
        float mapped = 1.0f - expf(-fbuffer[o] * 1.0f / (1 - exposure));
        ubuffer[o] = mapped * 255;
        return true;
    }
    unsigned char *getUCbuffer() { return ubuffer; }
    float *getFbuffer() { return fbuffer; }
    int getWidth() const { return width; }
    int getHeight() const { return height; }
    int getRenderCount() const { return curRenderCount; }
    void saveToFile(const std::string &) {}

  private:
    void Release() {
        delete[] ubuffer;
        delete[] fbuffer;
        ubuffer = nullptr;
        fbuffer = nullptr;
    }
    bool inside(int x, int y) const { return fbuffer && x >= 0 && y >= 0 && x < width && y < height; }
    size_t at(int x, int y, int c) const { return ((size_t)x + (size_t)y * width) * channals + c; }
    unsigned char *ubuffer;
    float *fbuffer;
    int width;
    int height;
    int channals;
    int curRenderCount;
};

#endif
