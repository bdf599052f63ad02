// material.h — host mirror of the reference's Material (material.h:23-68, material.cpp:82-162).
// The constructor parses the two text tables and builds tau, dE/dT, the three pdfs, their sums
// and the bulk conductivity exactly as the reference does; sampling (Material::Dist) and the
// per-step lookups (vel/tau/drawScatNext/scatter) run on the device, fed through
// mcb_upload_material (include/mcb.h).
#ifndef MCB_HOST_MATERIAL_H
#define MCB_HOST_MATERIAL_H
#include <iosfwd>
#include <string>
#include <vector>
#include "../../include/mcb.h"

#include "mc_types.h"
#include "phonon.h"

class Material {
    static const int nscat_ = 2;
    long np_, nw_;
    double T_, k_;
    std::vector<double> omega_, tau_, vel_;                 // (w,p) at [w + nw*p] like ArrayXXd(nw, np)
    std::vector<double> energyPdf_, fluxPdf_, scatPdf_;
    double energySum_, fluxSum_, scatSum_;
    std::string disp_, relax_;
    unsigned long uid_ = mcNextUid();                      // a copy shares the id: same tables
    std::string info() const;
public:
    Material();
    Material(const std::string& disp, const std::string& relax, double temp);

    double temp() const { return T_; }
    double cond() const { return k_; }
    long nw() const { return nw_; }
    long np() const { return np_; }
    unsigned long uid() const { return uid_; }
    double tau(long w, long p) const { return tau_.at((size_t)(w + nw_ * p)); }
    double vel(long w, long p) const { return vel_.at((size_t)(w + nw_ * p)); }
    double tau(const Phonon& phn) const { return tau(phn.prop().w(), phn.prop().p()); }      // material.cpp:174-184
    double vel(const Phonon& phn) const { return vel(phn.prop().w(), phn.prop().p()); }
    double energySum() const { return energySum_; }
    double fluxSum() const { return fluxSum_; }
    double scatSum() const { return scatSum_; }

    // the table view handed to mcb_upload_material (pointers stay valid for the Material's lifetime)
    mcb_material_desc desc() const;

    friend std::ostream& operator<<(std::ostream& os, const Material& mat);
};
#endif
