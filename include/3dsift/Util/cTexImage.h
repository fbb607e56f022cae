// 3dsift/Util/cTexImage.h — host view of one pyramid level, as exposed by CSIFT3D::GET_GSS() /
// GET_DOG() (reference: /root/reference/3DSIFT/Include/Util/cTexImage.h:8-60).  Public field and
// accessor names follow the reference; the data is a host copy of the device level (x fastest).
#ifndef S3D_FACADE_TEXIMAGE_H
#define S3D_FACADE_TEXIMAGE_H

#include <cstddef>
#include <vector>

class TexImage {
public:
    float* _Data = nullptr;   // points into `store`
    size_t _numsize = 0;
    int _nx = 0, _ny = 0, _nz = 0;
    float _s = 0.0f;          // scale of the level
    size_t _xs = 0, _ys = 0, _zs = 0;
    float _ux = 0.0f, _uy = 0.0f, _uz = 0.0f;

    TexImage() = default;
    TexImage(int width, int height, int depth) { SetImageSize(width, height, depth); }
    TexImage(const TexImage& o) { *this = o; }
    TexImage& operator=(const TexImage& o) {
        if (this != &o) {
            store = o.store;
            _numsize = o._numsize; _nx = o._nx; _ny = o._ny; _nz = o._nz; _s = o._s;
            _xs = o._xs; _ys = o._ys; _zs = o._zs; _ux = o._ux; _uy = o._uy; _uz = o._uz;
            _Data = store.empty() ? nullptr : store.data();
        }
        return *this;
    }

    void SetImageSize(int width, int height, int depth) {
        _nx = width; _ny = height; _nz = depth;
        _numsize = (size_t)width * height * depth;
        _xs = 1; _ys = (size_t)width; _zs = (size_t)width * height;
    }
    void SetImageScale(float scale) { _s = scale; }
    void SetImageUnit(float ux, float uy, float uz) { _ux = ux; _uy = uy; _uz = uz; }
    void MallocArrayMemory() { store.assign((size_t)_nx * _ny * _nz, 0.0f); _Data = store.data(); }

    float GetScale() const { return _s; }
    size_t GetXstride() const { return _xs; }
    size_t GetYstride() const { return _ys; }
    size_t GetZstride() const { return _zs; }
    float GetUnitX() const { return _ux; }
    float GetUnitY() const { return _uy; }
    float GetUnitZ() const { return _uz; }
    int GetDimX() const { return _nx; }
    int GetDimY() const { return _ny; }
    int GetDimZ() const { return _nz; }
    float GetImageDataWithIdx(int x, int y, int z) const { return _Data[(size_t)x * _xs + (size_t)y * _ys + (size_t)z * _zs]; }
    void SetImageDataWithIdx(float v, int x, int y, int z) { _Data[(size_t)x * _xs + (size_t)y * _ys + (size_t)z * _zs] = v; }

private:
    std::vector<float> store;
};

#endif
