// Stand-in for minkindr's <kindr/minimal/quat-transformation.h>, TEST INFRASTRUCTURE (oracle/_ref only).
// minkindr (ethz-asl, unpinned git URL in dependencies.rosinstall) is not vendored and not installed.
// Restated semantics: rotation = unit Eigen quaternion, rotate(v) = q * v, inverse = conjugate;
// T = (q, t): T * p = q.rotate(p) + t;  Ta * Tb = (qa * qb, ta + qa.rotate(tb));
// T^-1 = (q^-1, -(q^-1.rotate(t))).  No re-normalisation in products (constructors only CHECK the norm).
#ifndef NBV_REF_SHIM_KINDR_
#define NBV_REF_SHIM_KINDR_
#include <Eigen/Core>

namespace kindr {
namespace minimal {

template <class S>
class RotationQuaternionTemplate {
 public:
  typedef Eigen::Quaternion<S> Implementation;
  typedef Eigen::Matrix<S, 3, 1> Vector3;
  RotationQuaternionTemplate() : q_() {}
  RotationQuaternionTemplate(const Implementation& q) : q_(q) {}
  RotationQuaternionTemplate(S w, S x, S y, S z) : q_(w, x, y, z) {}
  const Implementation& toImplementation() const { return q_; }
  Vector3 rotate(const Vector3& v) const { return q_ * v; }
  Vector3 inverseRotate(const Vector3& v) const { return q_.conjugate() * v; }
  RotationQuaternionTemplate inverse() const { return RotationQuaternionTemplate(q_.conjugate()); }
  RotationQuaternionTemplate operator*(const RotationQuaternionTemplate& o) const { return RotationQuaternionTemplate(q_ * o.q_); }
  S w() const { return q_.w(); }
  S x() const { return q_.x(); }
  S y() const { return q_.y(); }
  S z() const { return q_.z(); }
 private:
  Implementation q_;
};

template <class S>
class AngleAxisTemplate {
 public:
  typedef Eigen::Matrix<S, 3, 1> Vector3;
  AngleAxisTemplate(S angle, const Vector3& axis) : angle_(angle), axis_(axis) {}
  // Rodrigues rotation; only reached in the reference's 360-degree lidar mode (32 corners), which is
  // undefined behaviour there and outside the parity contract.
  Vector3 rotate(const Vector3& v) const {
    const S c = std::cos(angle_), s = std::sin(angle_);
    return v * c + axis_.cross(v) * s + axis_ * (axis_.dot(v) * (S(1) - c));
  }
 private:
  S angle_;
  Vector3 axis_;
};

template <class S>
class QuatTransformationTemplate {
 public:
  typedef Eigen::Matrix<S, 3, 1> Position;
  typedef RotationQuaternionTemplate<S> Rotation;
  QuatTransformationTemplate() : q_(), t_() {}
  QuatTransformationTemplate(const Rotation& q, const Position& t) : q_(q), t_(t) {}
  QuatTransformationTemplate(const Position& t, const Rotation& q) : q_(q), t_(t) {}
  QuatTransformationTemplate(const typename Rotation::Implementation& q, const Position& t) : q_(q), t_(t) {}
  QuatTransformationTemplate(const Position& t, const typename Rotation::Implementation& q) : q_(q), t_(t) {}
  const Position& getPosition() const { return t_; }
  Position& getPosition() { return t_; }
  const Rotation& getRotation() const { return q_; }
  const typename Rotation::Implementation& getEigenQuaternion() const { return q_.toImplementation(); }
  Position transform(const Position& p) const { return q_.rotate(p) + t_; }
  Position operator*(const Position& p) const { return transform(p); }
  QuatTransformationTemplate operator*(const QuatTransformationTemplate& o) const {
    return QuatTransformationTemplate(q_ * o.q_, t_ + q_.rotate(o.t_));
  }
  QuatTransformationTemplate inverse() const { return QuatTransformationTemplate(q_.inverse(), -q_.inverseRotate(t_)); }
 private:
  Rotation q_;
  Position t_;
};

}  // namespace minimal
}  // namespace kindr
#endif
