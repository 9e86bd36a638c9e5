// Facade of ch4/v3/src/Object.h: shape descriptors.  The geometry tests themselves (inObject, lineIntersect) run on
// the device from a constant shape table built from these parameters (csrc/common.cuh, csrc/samplers.cuh).
#ifndef OBJECT_H
#define OBJECT_H
#include <ostream>
#include <stdexcept>
#include <string>
#include "Vec3.h"
#include "all.h"

class Object {
protected:
    std::string name = "Object";
    type_calc3 pos;
    type_calc phi = 0;

public:
    Object(type_calc3 pos_, type_calc phi_) : pos(pos_), phi(phi_) {}
    virtual ~Object() noexcept = default;
    void setPhi(type_calc p) noexcept { phi = p; }
    type_calc getPhi() noexcept { return phi; }
    type_calc3 getPos() const { return pos; }
    virtual void print(std::ostream& out) const { out << name << " pos: " << pos << " phi: " << phi; }
    virtual bool inObject(const type_calc3& x) const = 0;
    friend std::ostream& operator<<(std::ostream& out, const Object& o) { o.print(out); return out; }
};

class Sphere : public Object {
protected:
    type_calc radius, r_squared;

public:
    Sphere(type_calc3 pos_, type_calc phi_, type_calc r) : Object(pos_, phi_), radius(r), r_squared(r * r) { name = "Sphere"; }
    type_calc getRadius() const { return radius; }
    void print(std::ostream& out) const override { Object::print(out); out << " radius: " << radius; }
    bool inObject(const type_calc3& x) const override { type_calc3 r = x - pos; return r * r <= r_squared; }   // Object.cpp:111-115
};

class Rectangle : public Object {
protected:
    type_calc3 sides, half_sides, x_min, x_max;

public:
    Rectangle(type_calc3 pos_, type_calc phi_, type_calc3 sides_) : Object(pos_, phi_), sides(sides_) {
        name = "Rectangle";
        half_sides = sides_ * 0.5;                 // from the constructor argument, as the reference (Object.cpp:161,168)
        x_min = pos - half_sides; x_max = pos + half_sides;
    }
    Rectangle(type_calc3 pos_, type_calc phi_, type_calc3 sides_, type_calc3 /*orientation, unused in the reference too*/) : Rectangle(pos_, phi_, sides_) {}
    type_calc3 getSides() const { return sides; }
    void print(std::ostream& out) const override { Object::print(out); out << " sides: " << sides; }
    bool inObject(const type_calc3& x) const override {                                                       // Object.cpp:231-238
        type_calc3 t = abs(x - pos);
        for (int i = 0; i < 3; i++) if (t[i] > half_sides[i]) return false;
        return true;
    }
};
#endif
