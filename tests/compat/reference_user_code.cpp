// User code written the way a user of nickdou/montecarlocpp writes it -- same headers, same class and method names
// (phonon.h, field.h, material.h, domain.h, problem.h, random.h) -- compiled against the host mirror instead.
// tests/test_source_compat.py builds it with g++ and runs it: without a third argument only the host-side classes; with `gpu`
// also FieldProblem::solve (called from an OpenMP region like main.cpp:155-166) and Field::accumulate (device).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include "domain.h"
#include "field.h"
#include "material.h"
#include "phonon.h"
#include "problem.h"
#include "random.h"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: reference_user_code <disp> <relax> [gpu]\n"); return 2; }
    const bool gpu = argc > 3 && std::strcmp(argv[3], "gpu") == 0;

    // phonon.h:36-42, phonon.cpp:82-127
    Phonon phn(true, Phonon::Prop(3, 1), Vector3d(1e-8, 2e-8, 3e-8), Vector3d(3., 0., 4.));
    CHECK(phn.alive() && phn.sign() == 1 && phn.prop().w() == 3 && phn.prop().p() == 1);
    CHECK(std::fabs(phn.dir()(0) - 0.6) < 1e-15 && std::fabs(phn.dir()(2) - 0.8) < 1e-15);      // normalised on construction
    phn.scatNext(1e-7);
    phn.move(5e-8, 5000.);
    CHECK(std::fabs(phn.scatNext() - 5e-8) < 1e-22 && std::fabs(phn.time() - 1e-11) < 1e-25);
    CHECK(std::fabs(phn.pos()(0) - 4e-8) < 1e-22 && std::fabs(phn.pos()(2) - 7e-8) < 1e-22);
    phn.dir(Vector3d(0., 2., 0.), true);
    CHECK(phn.nscat() == 1 && phn.dir()(1) == 1. && phn.line().direction()(1) == 1.);
    bool refused = false;
    try { phn.scatNext(1.); } catch (const std::exception&) { refused = true; }                  // "Cannot reset scattering distance"
    CHECK(refused);
    TrkPhonon trk(phn);
    trk.pos(Vector3d(5e-8, 2e-8, 7e-8));
    CHECK(trk.trajectory().cols() == 2 && trk.trajectory().col(1)(0) == 5e-8);
    phn.kill();
    CHECK(!phn.alive());

    // material.h:54-57
    Material mat(argv[1], argv[2], 300.);
    Phonon probe(false, Phonon::Prop(mat.nw() / 2, 0), Vector3d(), Vector3d(1., 0., 0.));
    CHECK(mat.vel(probe) > 0. && mat.tau(probe) > 0. && probe.sign() == -1);
    CHECK(mat.cond() > 0. && mat.fluxSum() > 0. && mat.temp() == 300.);

    // domain.h, field.h:26-46, problem.h:151-155
    BulkDomain dom(Vector3d(1e-6, 1e-6, 1e-6), Vector3l(4, 2, 0), 1.);
    Field vol(1, &dom, CellVolF());
    CHECK(vol.data().rows() == 1 && vol.data().cols() == 8);
    CHECK(std::fabs(vol.data()(0, 3) - 1e-18 / 8.) < 1e-33);
    MultiProblem prob(&mat, &dom, 20000, 20);
    CHECK(prob.initSolution().rows() == 4 && prob.initSolution().cols() == 8);
    std::cout << mat << std::endl << dom << std::endl << prob << std::endl;
    if (!gpu) { std::printf("host-side ok\n"); return 0; }

    // main.cpp:146-179: every thread of the region enters solve() with its own generator, partial fields are summed
    ArrayXXd sol = prob.initSolution();
    Progress prog = prob.initProgress();
#pragma omp parallel num_threads(4)
    {
        Rng gen(1u);
        ArrayXXd partial = prob.solve(gen, &prog);
#pragma omp critical
        { sol += partial; }
    }
    CHECK(prog.count() == prob.nemit() && prog.esc() == 0);
    double q = 0.; for (long c = 0; c < sol.cols(); ++c) q += sol(1, c);
    CHECK(q > 0.);                                       // heat flows down the gradient
    // field.cpp:92-220 through the device: a segment across the 4 x 2 grid deposits its whole amount
    Field f(1, &dom);
    f.accumulate(dom.sdomPtrs().front(), Vector3d(1e-7, 1e-7, 5e-7), Vector3d(9e-7, 8e-7, 5e-7), VectorXd(1, 2.5));
    double tot = 0.; for (long c = 0; c < f.data().cols(); ++c) tot += f.data()(0, c);
    CHECK(std::fabs(tot - 2.5) < 1e-12);
    std::printf("gpu ok: %ld phonons, q_x sum %.6e\n", prog.count(), q);
    return 0;
}
