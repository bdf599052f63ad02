// material.cpp — table builder of the host Material mirror.  Follows material.cpp:23-45 (row-wise text
// extraction), :86-114 (file formats), :116-134 (tau model), :136-145 (dE/dT), :147-161 (pdfs, sums, k).
#include "material.h"
#include "constants.h"
#include "mc_types.h"
#include <fstream>
#include <iostream>
#include <sstream>

namespace {
// one text line per table row, `cols` whitespace-separated numbers per line
std::vector<double> readRows(std::istream& is, long rows, long cols, const char* what) {
    std::vector<double> a((size_t)(rows * cols));
    is >> std::ws;
    std::string line; long i = 0;
    for (; i < rows && std::getline(is, line); ++i) {
        std::stringstream ss(line);
        for (long j = 0; j < cols; ++j) {
            ss >> a[(size_t)(i * cols + j)];
            MC_ASSERT_MSG(ss, std::string("Array extraction failed: ") + what);
        }
    }
    MC_ASSERT_MSG(i == rows, std::string("Array extraction failed: ") + what);
    return a;
}
double total(const std::vector<double>& a) { double s = 0.; for (double x : a) s += x; return s; }
}

Material::Material() : np_(0), nw_(0), T_(0.), k_(0.), energySum_(0.), fluxSum_(0.), scatSum_(0.) {}

Material::Material(const std::string& disp, const std::string& relax, double temp)
    : np_(0), nw_(0), T_(temp), k_(0.), disp_(disp), relax_(relax) {
    std::ifstream dispFile(disp.c_str());
    MC_ASSERT_MSG(dispFile, "Error opening dispersion file");
    dispFile >> nw_ >> np_ >> std::ws;
    MC_ASSERT_MSG(nw_ > 0 && np_ > 0, "Invalid dispersion file");
    const long width = 2 + 2 * np_;
    const std::vector<double> table = readRows(dispFile, nw_, width, disp.c_str());

    const size_t n = (size_t)(nw_ * np_);
    omega_.resize((size_t)nw_); vel_.resize(n);
    std::vector<double> domega(n), dos(n);
    for (long w = 0; w < nw_; ++w) {
        const double* row = &table[(size_t)(w * width)];
        omega_[(size_t)w] = row[0];
        for (long p = 0; p < np_; ++p) {
            const size_t k = (size_t)(w + nw_ * p);
            domega[k] = row[1]; vel_[k] = row[2 + 2 * p]; dos[k] = row[3 + 2 * p];
        }
    }

    std::ifstream relaxFile(relax.c_str());
    MC_ASSERT_MSG(relaxFile, "Error opening relaxation time file");
    const std::vector<double> coeffs = readRows(relaxFile, np_, 4 * nscat_, relax.c_str());

    // tau^-1 = sum_j A w^a T^b exp(-c/T); a mechanism with A <= DBL_MIN is skipped
    tau_.assign(n, 0.);
    for (long p = 0; p < np_; ++p) {
        std::vector<double> rate((size_t)nw_, 0.);
        for (int j = 0; j < nscat_; ++j) {
            const double* c = &coeffs[(size_t)(p * 4 * nscat_ + 4 * j)];
            MC_ASSERT_MSG(c[0] >= 0., "Scattering times cannot be negative");
            if (c[0] <= Dbl::min()) continue;
            for (long w = 0; w < nw_; ++w)
                rate[(size_t)w] += (c[0] * std::pow(omega_[(size_t)w], c[1]) * std::pow(T_, c[2]) * std::exp(-c[3] / T_));
        }
        for (long w = 0; w < nw_; ++w) tau_[(size_t)(w + nw_ * p)] = 1. / rate[(size_t)w];
    }
    for (double t : tau_) MC_ASSERT_MSG(std::isfinite(t), "Scattering time model produced infinite values");

    // Bose-Einstein dE/dT per mode
    std::vector<double> dedT(n);
    for (long w = 0; w < nw_; ++w) {
        const double x = HBAR / (KB * T_) * omega_[(size_t)w];
        const double val = KB * (std::fabs(x) < Dbl::epsilon() ? 1. - x * x / 12. : std::pow(x / (2. * std::sinh(x / 2.)), 2));
        for (long p = 0; p < np_; ++p) dedT[(size_t)(w + nw_ * p)] = val;
    }

    energyPdf_.resize(n); fluxPdf_.resize(n); scatPdf_.resize(n);
    for (size_t i = 0; i < n; ++i) energyPdf_[i] = dedT[i] * dos[i] * domega[i];
    energySum_ = total(energyPdf_);
    for (size_t i = 0; i < n; ++i) fluxPdf_[i] = vel_[i] * energyPdf_[i];
    fluxSum_ = total(fluxPdf_);
    for (size_t i = 0; i < n; ++i) scatPdf_[i] = energyPdf_[i] / tau_[i];
    scatSum_ = total(scatPdf_);

    double ks = 0.;
    for (size_t i = 0; i < n; ++i) ks += tau_[i] * std::pow(vel_[i], 2) * energyPdf_[i];
    k_ = ks / 3.;
}

mcb_material_desc Material::desc() const {
    mcb_material_desc d;
    d.nw = nw_; d.np = np_; d.temp = T_;
    d.vel = vel_.data(); d.tau = tau_.data(); d.flux_pdf = fluxPdf_.data(); d.scat_pdf = scatPdf_.data();
    d.energy_sum = energySum_; d.flux_sum = fluxSum_; d.scat_sum = scatSum_;
    return d;
}

std::string Material::info() const {                          // material.cpp:233-241
    std::ostringstream ss;
    ss << "Material " << this << std::endl;
    ss << "  disp:  " << disp_ << std::endl;
    ss << "  relax: " << relax_ << std::endl;
    ss << "  temp:  " << T_;
    return ss.str();
}
std::ostream& operator<<(std::ostream& os, const Material& mat) { return os << mat.info(); }
