// Facade of ch4/v3/src/Field.h: a host-side mirror of a device node (or cell) field with the reference's
// element access f[i][j][k] and public ni/nj/nk.  The authoritative copy lives on the GPU; a mirror is refreshed
// lazily (one device->host copy) the first time it is read after the device changed it, which happens at output
// cadence in the reference loop (ch4/v3/src/main.cpp:264-284, Outputs.cpp:40-112).  Writes through operator[] mark
// the mirror as modified; the owner uploads it before the next device use.
#ifndef FIELD_H
#define FIELD_H
#include <functional>
#include <ostream>
#include <vector>
#include "Vec3.h"
#include "all.h"

template <class T>
class Field {
public:
    const int ni, nj, nk;
    const int nn[3];
    const int m_size;

    Field(int ni_, int nj_, int nk_) : ni(ni_), nj(nj_), nk(nk_), nn{ni_, nj_, nk_}, m_size(ni_ * nj_ * nk_), data((size_t)ni_ * nj_ * nk_) {}
    Field(const int n[3]) : Field(n[0], n[1], n[2]) {}
    Field(int3 n) : Field(n[0], n[1], n[2]) {}
    Field(const Field& o) : ni(o.ni), nj(o.nj), nk(o.nk), nn{o.ni, o.nj, o.nk}, m_size(o.m_size) { o.refresh(); data = o.data; }   // a plain host copy (e.g. `Field phi_0 = world.phi`)

    // [i][j][k] access through two light proxies over the flat (i*nj+j)*nk+k storage
    struct Row { T* p; T& operator[](int k) const { return p[k]; } };
    struct Plane { T* p; int nk; Row operator[](int j) const { return Row{p + (size_t)j * nk}; } };
    struct CRow { const T* p; const T& operator[](int k) const { return p[k]; } };
    struct CPlane { const T* p; int nk; CRow operator[](int j) const { return CRow{p + (size_t)j * nk}; } };
    Plane operator[](int i) { refresh(); host_modified = true; return Plane{data.data() + (size_t)i * nj * nk, nk}; }
    CPlane operator[](int i) const { refresh(); return CPlane{data.data() + (size_t)i * nj * nk, nk}; }

    int U(int i, int j, int k) const { return k * ni * nj + j * ni + i; }      // the solver's flat index (Field.h:263-265)
    int size() const { return m_size; }
    void clear() { (*this) = T{}; }
    Field& operator=(const T v) { std::fill(data.begin(), data.end(), v); stale = false; host_modified = true; return *this; }
    Field& operator=(const Field& o) { o.refresh(); data = o.data; stale = false; host_modified = true; return *this; }

    // ---- device coupling (used by World / Species)
    T* raw() { return data.data(); }
    const T* raw() const { return data.data(); }
    void bind(std::function<void(T*)> fetch_) { fetch = std::move(fetch_); }
    void invalidate() const { stale = true; }                  // the device copy changed
    bool takeHostModified() { bool m = host_modified; host_modified = false; return m; }
    void refresh() const { if (stale && fetch) { stale = false; fetch(const_cast<T*>(data.data())); } }

private:
    std::vector<T> data;
    std::function<void(T*)> fetch;
    mutable bool stale = false;
    bool host_modified = false;
};

template <class T>
std::ostream& operator<<(std::ostream& out, const Field<T>& f) {      // VTK-style dump: k slowest, i fastest (Field.h print order)
    for (int k = 0; k < f.nk; k++, out << "\n")
        for (int j = 0; j < f.nj; j++)
            for (int i = 0; i < f.ni; i++) out << f[i][j][k] << " ";
    return out;
}
#endif
