// Oracle driver: runs the UNMODIFIED reference (headers and sources compiled where they lie
// under /root/reference, see oracle/Makefile) through its own public API and dumps every
// node's operators/vectors, or times the three stages.  Test infrastructure only: nothing in
// the product links or executes this.  API use mirrors examples/elliptic-single/main.cpp:104-217.
//
//   ref_driver --problem poisson|helmholtz|varcoef --solver fishpack|fivepoint
//              --min-level a --max-level b --nx M --domain xl xu yl yu --threshold t
//              [--homogeneous 0|1] [--cache 0|1] [--nsolves k] [--dump file] [--ops 0|1]
//              [--refine-box x0 x1 y0 y1]   (refine inside a box instead of |f| > threshold)
//              [--lambda v]                 (with --problem helmholtz: constant lambda = v instead of -1)
//              [--dump-root-only 1]         (dump only the root's T, S, h, w and every leaf's u: parity at bench scale)
#include <EllipticForest.hpp>
#include <Patches/FiniteVolume/FiniteVolume.hpp>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>

using namespace EllipticForest;
using PatchT = FiniteVolumePatch;
using NodeT = Node<PatchT>;

static FILE* dumpf = nullptr;
static void rec(const std::string& name, const std::vector<long>& dims, const double* data) {
    if (!dumpf) return;
    int nl = (int)name.size(); int nd = (int)dims.size();
    fwrite(&nl, 4, 1, dumpf); fwrite(name.data(), 1, nl, dumpf); fwrite(&nd, 4, 1, dumpf);
    long tot = 1; for (long d : dims) { fwrite(&d, 8, 1, dumpf); tot *= d; }
    fwrite(data, 8, tot, dumpf);
}
static void recMat(const std::string& name, Matrix<double>& A) { if (A.nRows() * A.nCols() > 0) rec(name, {(long)A.nRows(), (long)A.nCols()}, A.dataPointer()); }
static void recVec(const std::string& name, Vector<double>& v) { if (v.size() > 0) rec(name, {(long)v.size()}, v.dataPointer()); }

struct Problem {
    std::string name; double lambda0;
    double u(double x, double y) const { return sin(x) + sin(y); }
    double alpha(double, double) const { return 1.0; }
    double beta(double x, double y) const { return name == "varcoef" ? 1.0 + 0.5 * sin(x) * cos(y) : 1.0; }
    double lambda(double x, double y) const { return name == "varcoef" ? -(1.0 + 0.5 * cos(x) * cos(y)) : lambda0; }
    double f(double x, double y) const {
        if (name == "varcoef") {
            double bx = 0.5 * cos(x) * cos(y), by = -0.5 * sin(x) * sin(y);
            return bx * cos(x) + by * cos(y) - beta(x, y) * u(x, y) + lambda(x, y) * u(x, y);
        }
        return (lambda0 - 1.0) * u(x, y);
    }
};

int main(int argc, char** argv) {
    Problem P{"poisson", 0.0};
    std::string solver_name = "fishpack", dump;
    int min_level = 0, max_level = 2, nx = 8, nsolves = 1; bool homogeneous = false, cache = false, ops = true, root_only = false;
    double xl = -10, xu = 10, yl = -10, yu = 10, threshold = 1.2;
    bool use_box = false; double rb[4] = {0, 0, 0, 0};
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() { return std::string(argv[++i]); };
        if (a == "--problem") { P.name = next(); P.lambda0 = (P.name == "helmholtz") ? -1.0 : 0.0; }
        else if (a == "--lambda") P.lambda0 = std::stod(next());   // after --problem helmholtz: constant lambda (> 0: indefinite)
        else if (a == "--solver") solver_name = next();
        else if (a == "--min-level") min_level = std::stoi(next());
        else if (a == "--max-level") max_level = std::stoi(next());
        else if (a == "--nx") nx = std::stoi(next());
        else if (a == "--threshold") threshold = std::stod(next());
        else if (a == "--homogeneous") homogeneous = std::stoi(next());
        else if (a == "--cache") cache = std::stoi(next());
        else if (a == "--nsolves") nsolves = std::stoi(next());
        else if (a == "--ops") ops = std::stoi(next());
        else if (a == "--dump") dump = next();
        else if (a == "--dump-root-only") root_only = std::stoi(next());
        else if (a == "--refine-box") { use_box = true; for (int k = 0; k < 4; k++) rb[k] = std::stod(next()); }
        else if (a == "--domain") { xl = std::stod(next()); xu = std::stod(next()); yl = std::stod(next()); yu = std::stod(next()); }
        else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
    }
    int fargc = 1; char** fargv = argv;
    EllipticForestApp app(&fargc, &fargv);
    app.options.setOption("cache-operators", cache);
    app.options.setOption("homogeneous-rhs", homogeneous);

    FiniteVolumeGrid grid(MPI_COMM_WORLD, nx, xl, xu, nx, yl, yu);
    FiniteVolumePatch root_patch(MPI_COMM_WORLD, grid);
    FiniteVolumeNodeFactory node_factory(MPI_COMM_WORLD);
    Mesh<FiniteVolumePatch> mesh{};
    auto t0 = std::chrono::steady_clock::now();
    mesh.refineByFunction([&](double x, double y) { return use_box ? (x > rb[0] && x < rb[1] && y > rb[2] && y < rb[3]) : fabs(-(sin(x) + sin(y))) > threshold; },
                          threshold, min_level, max_level, root_patch, node_factory);
    double t_mesh = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    FiniteVolumeSolver solver{};
    solver.solver_type = solver_name == "fishpack" ? FiniteVolumeSolverType::FISHPACK90 : FiniteVolumeSolverType::FivePointStencil;
    solver.alpha_function = [&](double x, double y) { return P.alpha(x, y); };
    solver.beta_function = [&](double x, double y) { return P.beta(x, y); };
    solver.lambda_function = [&](double x, double y) { return P.lambda(x, y); };

    HPSAlgorithm<FiniteVolumeGrid, FiniteVolumeSolver, FiniteVolumePatch, double> HPS(MPI_COMM_WORLD, mesh, solver);
    if (!dump.empty()) dumpf = fopen(dump.c_str(), "wb");

    // traversal orders (ordering contract) + node table
    std::string post, pre; long nleaves = 0;
    mesh.quadtree.traversePostOrder([&](NodeT* n) { post += n->path + (n->leaf ? "L" : "P") + ";"; return 1; });
    mesh.quadtree.traversePreOrder([&](NodeT* n) { pre += n->path + ";"; if (n->leaf) nleaves++; return 1; });
    if (dumpf) {
        std::vector<double> pb(post.begin(), post.end()), qb(pre.begin(), pre.end());
        rec("order/post", {(long)pb.size()}, pb.data()); rec("order/pre", {(long)qb.size()}, qb.data());
        mesh.quadtree.traversePreOrder([&](NodeT* n) {
            if (root_only && n->path != "0") return 1;
            auto& g = n->data.grid();
            double box[6] = {g.xLower(), g.xUpper(), g.yLower(), g.yUpper(), (double)g.nx(), (double)n->level};
            rec("grid0/" + n->path, {6}, box); return 1; });
    }

    HPS.setupStage();
    HPS.buildStage();
    double t_build = app.timers["build-stage"].time();
    if (dumpf) mesh.quadtree.traversePostOrder([&](NodeT* n) {
        if (root_only && n->path != "0") return 1;
        auto& p = n->data; std::string k = "build/" + n->path + "/";
        double meta[3] = {(double)p.n_coarsens, (double)p.grid().nx(), (double)n->leaf};
        rec(k + "meta", {3}, meta);
        if (ops) { recMat(k + "T", p.matrixT()); recMat(k + "S", p.matrixS()); if (!root_only) { recMat(k + "X", p.matrixX()); recMat(k + "H", p.matrixH()); } }
        return 1; });

    double t_up = 0, t_solve = 0;
    for (int s = 0; s < nsolves; s++) {
        double scale = 1.0 + s / 100.0;
        HPS.upwardsStage([&](double x, double y) { return scale * P.f(x, y); });
        t_up += app.timers["upwards-stage"].time();
        if (dumpf && s == nsolves - 1) mesh.quadtree.traversePostOrder([&](NodeT* n) {
            if (root_only && n->path != "0") return 1;
            auto& p = n->data; std::string k = "up/" + n->path + "/";
            recVec(k + "h", p.vectorH()); recVec(k + "w", p.vectorW()); if (n->leaf) recVec(k + "f", p.vectorF());
            return 1; });
        HPS.solveStage([&](int side, double x, double y, double* a, double* b) { *a = 1.0; *b = 0.0; return scale * P.u(x, y); });
        t_solve += app.timers["solve-stage"].time();
        if (dumpf && s == nsolves - 1) mesh.quadtree.traversePostOrder([&](NodeT* n) {
            auto& p = n->data; std::string k = "solve/" + n->path + "/";
            if (!root_only || n->path == "0") recVec(k + "g", p.vectorG());
            if (n->leaf) recVec(k + "u", p.vectorU());
            return 1; });
    }
    // error vs manufactured solution on the last solve
    double scale = 1.0 + (nsolves - 1) / 100.0, emax = 0;
    mesh.quadtree.traversePreOrder([&](NodeT* n) {
        if (!n->leaf) return 1;
        auto& g = n->data.grid(); auto& u = n->data.vectorU();
        for (int i = 0; i < (int)g.nx(); i++) for (int j = 0; j < (int)g.ny(); j++)
            emax = fmax(emax, fabs(u[j + i * g.ny()] - scale * P.u(g(0, i), g(1, j))));
        return 1; });
    if (dumpf) fclose(dumpf);
    printf("REF_RESULT {\"leaves\": %ld, \"dofs\": %ld, \"mesh_s\": %.6f, \"build_s\": %.6f, \"upwards_s\": %.6f, \"solve_s\": %.6f, \"nsolves\": %d, \"linf_error\": %.6e}\n",
           nleaves, nleaves * nx * nx, t_mesh, t_build, t_up / nsolves, t_solve / nsolves, nsolves, emax);
    fflush(stdout);
    _exit(0);  // skip the app destructor's timer dump / MPI finalize chatter
}
