// phonon.h — host mirror of Phonon / TrkPhonon (phonon.h:20-89, phonon.cpp:21-170) so that user code written against the
// reference's headers compiles against this mirror.  On the accelerated path a phonon lives on the device as one slot of the
// warp-tiled state (pos, dir, scatNext, packed integers: mcb_device.cuh); these value classes carry the same state and the
// same setter semantics on the host (normalising dir(), the scatNext bookkeeping of move(), the polyline of TrkPhonon).
#ifndef MCB_HOST_PHONON_H
#define MCB_HOST_PHONON_H
#include <vector>
#include "constants.h"
#include "mc_types.h"

// the one member of Eigen::ParametrizedLine<double, 3> the reference uses besides construction (phonon.cpp:104)
struct ParametrizedLine3 {
    Vector3d o, d;
    ParametrizedLine3() {}
    ParametrizedLine3(const Vector3d& origin, const Vector3d& direction) : o(origin), d(direction) {}
    const Vector3d& origin() const { return o; }
    const Vector3d& direction() const { return d; }
    Vector3d pointAt(double t) const { return o + t * d; }
};

class Phonon {
public:
    class Prop {
        long w_, p_;
    public:
        Prop() : w_(0), p_(0) {}
        Prop(long omega, long pol) : w_(omega), p_(pol) {}
        long w() const { return w_; }
        long p() const { return p_; }
    };
private:
    bool alive_, sign_;
    Prop prop_;
    Vector3d pos_, dir_;
    ParametrizedLine3 line_;
    double time_, scatNext_;
    long nscat_;
public:
    Phonon() : alive_(false), sign_(false), time_(0.), scatNext_(0.), nscat_(0) {}
    Phonon(bool sign, const Prop& prop, const Vector3d& pos, const Vector3d& dir)
        : alive_(true), sign_(sign), prop_(prop), pos_(pos), dir_(dir.normalized()), line_(pos, dir.normalized()),
          time_(0.), scatNext_(0.), nscat_(0) {}
    virtual ~Phonon() {}

    bool alive() const { return alive_; }
    void kill() { alive_ = false; }
    int sign() const { return sign_ ? 1 : -1; }
    const Prop& prop() const { return prop_; }
    void prop(const Prop& p) { prop_ = p; }

    virtual const Vector3d& pos() const { return pos_; }
    const Vector3d& dir() const { return dir_; }
    const ParametrizedLine3& line() const { return line_; }

    virtual void pos(const Vector3d& newPos) { pos_ = newPos; line_ = ParametrizedLine3(pos_, dir_); }
    void dir(const Vector3d& newDir, bool scatter) {               // normalises on every set; counts a scattering event
        dir_ = newDir.normalized();
        line_ = ParametrizedLine3(pos_, dir_);
        if (scatter) nscat_++;
    }
    void move(double distance, double vel) {
        MC_ASSERT_MSG(distance <= scatNext_, "Movement distance too large");
        scatNext_ -= distance;
        if (scatNext_ < Dbl::min()) scatNext_ = 0.;
        MC_ASSERT_MSG(vel > Dbl::min(), "Velocity must be nonzero");
        time_ += distance / vel;
        pos(line_.pointAt(distance));
    }
    double time() const { return time_; }
    double scatNext() const { return scatNext_; }
    long nscat() const { return nscat_; }
    void scatNext(double distance) {
        MC_ASSERT_MSG(scatNext_ == 0., "Cannot reset scattering distance");
        MC_ASSERT_MSG(distance > 0., "Scattering distance must be positive");
        scatNext_ = distance;
    }
};

// records every position it is given (phonon.cpp:129-170): the polyline the `traj` / `check` modes print
class TrkPhonon : public Phonon {
    std::vector<Vector3d> traj_;
public:
    TrkPhonon() : Phonon() {}
    TrkPhonon(bool sign, const Prop& prop, const Vector3d& pos, const Vector3d& dir) : Phonon(sign, prop, pos, dir), traj_(1, pos) {}
    TrkPhonon(const Phonon& phn) : Phonon(phn), traj_(1, phn.pos()) {}
    ~TrkPhonon() {}
    const Vector3d& pos() const { return Phonon::pos(); }
    void pos(const Vector3d& newPos) { Phonon::pos(newPos); traj_.push_back(newPos); }
    Matrix3Xd trajectory() const { Matrix3Xd m; m.c = traj_; return m; }
};
#endif
