// EllipticForestB200.hpp - reference-side binding of the B200 HPS path.
//
// A maintainer of EllipticForest adds this ONE header next to the library's own (it includes the
// reference's public headers unchanged) and links libefgpu.so.  It derives from the reference's
//     HPSAlgorithm<FiniteVolumeGrid, FiniteVolumeSolver, FiniteVolumePatch, double>
// (src/HPSAlgorithm.hpp:26-82) and overrides the four virtual stages (setupStage :91, buildStage
// :120, upwardsStage :178/:225, solveStage :291/:343) so that an existing driver changes one type
// name and nothing else:
//
//     EllipticForest::HPSAlgorithmB200 HPS(MPI_COMM_WORLD, mesh, solver);   // was HPSAlgorithm<...>
//     HPS.setupStage(); HPS.buildStage(); HPS.upwardsStage(f); HPS.solveStage(bc);
//     mesh.quadtree.traversePreOrder(... patch.vectorU() ...)               // unchanged
//
// Mesh<>, Quadtree<> (p4est ordering), FiniteVolumeSolver (solver_type, alpha/beta/lambda
// functions), FiniteVolumePatch and the app options "cache-operators" / "homogeneous-rhs" and the
// four stage timers stay the reference's.  All arithmetic of the path runs behind the C-ABI of
// include/efgpu.h; nothing here falls back to the CPU implementation.
//
// This header is host-only C++20 and contains no CUDA.  oracle/dropin_driver.cpp compiles it
// against the unmodified reference sources and checks it against the reference's own stages.
#ifndef ELLIPTIC_FOREST_B200_HPP_
#define ELLIPTIC_FOREST_B200_HPP_

#include <EllipticForest.hpp>
#include <Patches/FiniteVolume/FiniteVolume.hpp>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <exception>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "efgpu.h"

namespace EllipticForest {

class HPSAlgorithmB200 : public HPSAlgorithm<FiniteVolumeGrid, FiniteVolumeSolver, FiniteVolumePatch, double> {
public:
    using Base = HPSAlgorithm<FiniteVolumeGrid, FiniteVolumeSolver, FiniteVolumePatch, double>;
    using PatchT = FiniteVolumePatch;
    using NodeT = Node<FiniteVolumePatch>;

    int device = 0;
    // Parity/debug: also copy T,S,X,H (build), h,w (upwards) and g (solve) of EVERY node back into
    // the patches, as the reference leaves them.  Off by default: only leaf vectorU()/vectorF(),
    // grid() and n_coarsens are written, which is all the reference's drivers read.
    bool copy_back_operators = false;
    // The reference calls rhs_function / alpha / beta / lambda point by point on the calling thread
    // (HPSAlgorithm.hpp:241-249, FiniteVolumeSolver.cpp:63-79); at 1e7+ cells that loop, not the stages, is the wall
    // clock.  > 1: the leaves are sampled in contiguous blocks on this many host threads - only for callbacks that are
    // thread-safe, hence opt-in.  The values and where they are stored do not depend on it.
    int sampling_threads = 1;
    // Host threads for the binding's own per-leaf copies that call no user code and no MPI (u into every leaf's vectorU): at
    // 1e7 cells these loops over the reference's containers cost more than the stages on the device (bench.py:
    // e2e_cpp_binding).  The values written do not depend on it.
    int copy_threads = 8;

    HPSAlgorithmB200(MPI::Communicator comm, Mesh<PatchT>& mesh, FiniteVolumeSolver& solver, int device = 0)
        : Base(comm, mesh, solver), device(device) {}
    HPSAlgorithmB200(const HPSAlgorithmB200&) = delete;
    HPSAlgorithmB200& operator=(const HPSAlgorithmB200&) = delete;
    ~HPSAlgorithmB200() { if (h_) efgpu_destroy(h_); }

    efgpu_handle* handle() { return h_; }
    const std::vector<NodeT*>& nodes() const { return nodes_; }     // p4est pre-order, node id = index

    // ---- setupStage: flatten the reference's quadtree in ITS OWN traversal order ---------------
    void setupStage() override {
        EllipticForestApp& app = EllipticForestApp::getInstance();
        app.addTimer("setup-stage");
        app.timers["setup-stage"].start();
        flatten_();
        efgpu_tree_desc d;
        d.n_nodes = (int32_t)nodes_.size();
        d.nx = nx_;
        d.level = level_.data(); d.child = child_.data(); d.box = box_.data();
        if (h_) { efgpu_destroy(h_); h_ = nullptr; }
        check_(efgpu_create(&d, device, &h_), "efgpu_create");
        app.timers["setup-stage"].stop();
    }

    // ---- buildStage (HPSAlgorithm.hpp:120-161) -------------------------------------------------
    void buildStage() override {
        EllipticForestApp& app = EllipticForestApp::getInstance();
        if (!h_) setupStage();
        app.addTimer("build-stage");
        app.timers["build-stage"].start();
        if (this->patch_solver.solver_type == FiniteVolumeSolverType::FISHPACK90) {
            // FISHPACK branch ignores alpha/beta and evaluates lambda at the origin (FiniteVolumeSolver.cpp:254)
            check_(efgpu_set_leaf_constant(h_, this->patch_solver.lambda_function(0.0, 0.0)), "efgpu_set_leaf_constant");
        } else {
            sample_coefficients_();
        }
        check_(efgpu_build(h_, flags_()), "efgpu_build");
        // mergePatch_ (:1004-1009) and coarsen_ (:736): merged grids and n_coarsens are visible to callers
        for (size_t i = nodes_.size(); i-- > 0;) {     // (serial: the grid constructor queries the MPI communicator)
            int size = 0, nco = 0, leaf = 0, li = 0;
            efgpu_node_info(h_, (int)i, &size, &nco, &leaf, &li);
            PatchT& p = nodes_[i]->data;
            if (!leaf) {
                const double* b = &box_[4 * i];
                p.grid() = FiniteVolumeGrid(MPI_COMM_SELF, size, b[0], b[1], size, b[2], b[3]);
            }
            p.n_coarsens = nco;
        }
        if (copy_back_operators) {
            for (size_t i = 0; i < nodes_.size(); i++) {
                PatchT& p = nodes_[i]->data;
                fetch_matrix_((int)i, EFGPU_OP_T, p.matrixT());
                if (!nodes_[i]->leaf) {
                    fetch_matrix_((int)i, EFGPU_OP_S, p.matrixS());
                    fetch_matrix_((int)i, EFGPU_OP_H, p.matrixH());
                    if (keep_x) fetch_matrix_((int)i, EFGPU_OP_X, p.matrixX());
                }
            }
        }
        this->isBuilt = true;
        app.timers["build-stage"].stop();
    }
    bool keep_x = false;   // with copy_back_operators: also retain and copy X (the product stores X^-1 only)

    // ---- upwardsStage(f(x,y)) (HPSAlgorithm.hpp:225-272) ---------------------------------------
    void upwardsStage(std::function<double(double, double)> rhs_function) override {
        EllipticForestApp& app = EllipticForestApp::getInstance();
        app.addTimer("upwards-stage");
        app.timers["upwards-stage"].start();
        const size_t cells = (size_t)nx_ * nx_;
        f_.resize(leaves_.size() * cells);
        for_leaves_([&](size_t l) {
            PatchT& patch = nodes_[leaves_[l]]->data;
            FiniteVolumeGrid& grid = patch.grid();
            if ((size_t)patch.vectorF().size() != cells) patch.vectorF() = Vector<double>(cells);
            for (int i = 0; i < nx_; i++) {
                const double x = grid(0, i);
                for (int j = 0; j < nx_; j++) {
                    const double v = rhs_function(x, grid(1, j));
                    patch.vectorF()[j + i * nx_] = v;                   // :246-247
                    f_[l * cells + j + (size_t)i * nx_] = v;
                }
            }
        });
        run_upwards_();
        app.timers["upwards-stage"].stop();
    }

    // ---- upwardsStage(fn(Patch&)) (HPSAlgorithm.hpp:178-216) -----------------------------------
    void upwardsStage(std::function<void(PatchT& leafPatch)> rhs_patch_function) override {
        EllipticForestApp& app = EllipticForestApp::getInstance();
        app.addTimer("upwards-stage");
        app.timers["upwards-stage"].start();
        const size_t cells = (size_t)nx_ * nx_;
        f_.resize(leaves_.size() * cells);
        for (size_t l = 0; l < leaves_.size(); l++) {
            PatchT& patch = nodes_[leaves_[l]]->data;
            rhs_patch_function(patch);
            if ((size_t)patch.vectorF().size() != cells) throw std::invalid_argument("[EllipticForest::HPSAlgorithmB200::upwardsStage] vectorF has the wrong size");
            for (size_t c = 0; c < cells; c++) f_[l * cells + c] = patch.vectorF()[c];
        }
        run_upwards_();
        app.timers["upwards-stage"].stop();
    }

    // ---- solveStage(fn(Patch& root)) (HPSAlgorithm.hpp:291-324) --------------------------------
    void solveStage(std::function<void(PatchT& rootPatch)> boundary_data_function) override {
        EllipticForestApp& app = EllipticForestApp::getInstance();
        app.addTimer("solve-stage");
        app.timers["solve-stage"].start();
        PatchT& root = this->mesh.quadtree.root();
        boundary_data_function(root);
        int size = 0; efgpu_node_info(h_, 0, &size, nullptr, nullptr, nullptr);
        if (root.vectorG().size() != 4 * size) throw std::invalid_argument("[EllipticForest::HPSAlgorithmB200::solveStage] root vectorG has the wrong size");
        u_.resize(f_size_());
        check_(efgpu_solve_dirichlet(h_, root.vectorG().dataPointer(), flags_(), u_.data()), "efgpu_solve_dirichlet");
        scatter_solution_();
        app.timers["solve-stage"].stop();
    }

    // ---- solveStage(fn(side,x,y,*a,*b)) (HPSAlgorithm.hpp:343-445) -----------------------------
    void solveStage(std::function<double(int side, double x, double y, double* a, double* b)> boundary_analytical_function) override {
        EllipticForestApp& app = EllipticForestApp::getInstance();
        app.addTimer("solve-stage");
        app.timers["solve-stage"].start();
        PatchT& rootPatch = this->mesh.quadtree.root();
        FiniteVolumeGrid& rootGrid = rootPatch.grid();     // merged grid, set by buildStage
        const int M = rootGrid.nx();
        std::vector<double> r(4 * (size_t)M), a(4 * (size_t)M), b(4 * (size_t)M);
        for (int n = 0; n < 4; n++)                        // sampling points exactly as :375-400
            for (int i = 0; i < M; i++) {
                double x, y;
                if (n == 0) { x = rootGrid.xLower(); y = rootGrid(1, i); }
                else if (n == 1) { x = rootGrid.xUpper(); y = rootGrid(1, i); }
                else if (n == 2) { x = rootGrid(0, i); y = rootGrid.yLower(); }
                else { x = rootGrid(0, i); y = rootGrid.yUpper(); }
                double av, bv;
                r[(size_t)n * M + i] = boundary_analytical_function(n, x, y, &av, &bv);
                a[(size_t)n * M + i] = av; b[(size_t)n * M + i] = bv;
            }
        u_.resize(f_size_());
        check_(efgpu_solve_robin(h_, a.data(), b.data(), r.data(), flags_(), u_.data()), "efgpu_solve_robin");
        scatter_solution_();
        app.timers["solve-stage"].stop();
    }

private:
    efgpu_handle* h_ = nullptr;
    int nx_ = 0;
    std::vector<NodeT*> nodes_;
    std::vector<int> leaves_;
    std::vector<int32_t> level_, child_;
    std::vector<double> box_, f_, u_;

    size_t f_size_() const { return leaves_.size() * (size_t)nx_ * nx_; }

    unsigned flags_() {
        EllipticForestApp& app = EllipticForestApp::getInstance();
        unsigned f = 0;
        if (std::get<bool>(app.options["cache-operators"])) f |= EFGPU_CACHE_OPERATORS;   // bad_variant_access if unset, as the reference (:134)
        if (std::get<bool>(app.options["homogeneous-rhs"])) f |= EFGPU_HOMOGENEOUS_RHS;
        if (copy_back_operators && keep_x) f |= EFGPU_KEEP_X;
        return f;
    }

    void check_(int status, const char* what) {
        if (status == EFGPU_OK) return;
        const std::string msg = std::string("[EllipticForest::HPSAlgorithmB200] ") + what + ": " + efgpu_last_error(h_);
        switch (status) {
            case EFGPU_ERR_BAD_ARG: case EFGPU_ERR_BAD_SHAPE: throw std::invalid_argument(msg);   // reference: HPSAlgorithm.hpp:761
            case EFGPU_ERR_SINGULAR: std::cerr << msg << " (continuing, as the reference does on LAPACK INFO > 0)" << std::endl; return;   // Matrix.hpp:898,947
            default: throw std::runtime_error(msg);
        }
    }

    // node table = what traversePreOrder visits (Quadtree.hpp:236-260); children located through the
    // reference's own path keys ("0" + child ids, P4est.cpp:35-44)
    void flatten_() {
        nodes_.clear(); leaves_.clear();
        std::map<std::string, int> id;
        this->mesh.quadtree.traversePreOrder([&](NodeT* n) {
            id[n->path] = (int)nodes_.size();
            nodes_.push_back(n);
            return 1;
        });
        const size_t nn = nodes_.size();
        level_.assign(nn, 0); child_.assign(4 * nn, -1); box_.assign(4 * nn, 0.0);
        nx_ = 0;
        for (size_t i = 0; i < nn; i++) {
            NodeT* n = nodes_[i];
            level_[i] = n->level;
            FiniteVolumeGrid& g = n->data.grid();
            box_[4 * i + 0] = g.xLower(); box_[4 * i + 1] = g.xUpper(); box_[4 * i + 2] = g.yLower(); box_[4 * i + 3] = g.yUpper();
            if (n->leaf) {
                if (g.nx() != g.ny()) throw std::invalid_argument("[EllipticForest::HPSAlgorithmB200] square patches only (nx == ny)");
                if (nx_ == 0) nx_ = g.nx();
                if (nx_ != (int)g.nx()) throw std::invalid_argument("[EllipticForest::HPSAlgorithmB200] all leaf patches must have the same nx");
                leaves_.push_back((int)i);
            } else {
                for (int c = 0; c < 4; c++) {
                    auto it = id.find(n->path + std::to_string(c));
                    if (it == id.end()) throw std::invalid_argument("[EllipticForest::HPSAlgorithmB200] node " + n->path + " misses child " + std::to_string(c) + " (rank-shared trees: use the sharded entry)");
                    child_[4 * i + c] = it->second;
                }
            }
        }
    }

    // body(l) for every leaf l, serially or in contiguous blocks on `sampling_threads` host threads
    template <class F>
    void for_leaves_(F&& body) { parallel_for_(leaves_.size(), sampling_threads, body); }

    // body(i) for i in [0, n), serially or in contiguous blocks on `threads` host threads
    template <class F>
    void parallel_for_(size_t n, int threads, F&& body) {
        const size_t nt = std::min<size_t>((size_t)std::max(1, threads), n);
        if (nt <= 1) { for (size_t l = 0; l < n; l++) body(l); return; }
        std::vector<std::thread> pool;
        std::exception_ptr err;
        std::mutex guard;
        for (size_t t = 0; t < nt; t++)
            pool.emplace_back([&, t] {
                try { for (size_t l = n * t / nt; l < n * (t + 1) / nt; l++) body(l); }
                catch (...) { std::lock_guard<std::mutex> lock(guard); if (!err) err = std::current_exception(); }
            });
        for (auto& th : pool) th.join();
        if (err) std::rethrow_exception(err);
    }

    // alpha, lambda at cell centres; beta at face midpoints: the sampling points of FiniteVolumeSolver.cpp:63-79
    void sample_coefficients_() {
        const size_t cells = (size_t)nx_ * nx_, tot = leaves_.size() * cells;
        std::vector<double> al(tot), bw(tot), be(tot), bs(tot), bn(tot), la(tot);
        FiniteVolumeSolver& s = this->patch_solver;
        for_leaves_([&](size_t l) {
            FiniteVolumeGrid& grid = nodes_[leaves_[l]]->data.grid();
            const double dx = grid.dx(), dy = grid.dy();
            for (int i = 0; i < nx_; i++)
                for (int j = 0; j < nx_; j++) {
                    const double xi = grid(0, i), yj = grid(1, j);
                    const size_t k = l * cells + j + (size_t)i * nx_;
                    al[k] = s.alpha_function(xi, yj);
                    be[k] = s.beta_function(xi + dx / 2.0, yj);
                    bw[k] = s.beta_function(xi - dx / 2.0, yj);
                    bn[k] = s.beta_function(xi, yj + dy / 2.0);
                    bs[k] = s.beta_function(xi, yj - dy / 2.0);
                    la[k] = s.lambda_function(xi, yj);
                }
        });
        check_(efgpu_set_leaf_variable(h_, al.data(), bw.data(), be.data(), bs.data(), bn.data(), la.data()), "efgpu_set_leaf_variable");
    }

    void run_upwards_() {
        check_(efgpu_upwards(h_, f_.data(), 1.0, flags_()), "efgpu_upwards");
        if (!copy_back_operators) return;
        for (size_t i = 0; i < nodes_.size(); i++) {
            fetch_vector_((int)i, EFGPU_VEC_H, nodes_[i]->data.vectorH());
            if (!nodes_[i]->leaf) fetch_vector_((int)i, EFGPU_VEC_W, nodes_[i]->data.vectorW());
        }
    }

    void scatter_solution_() {
        const size_t cells = (size_t)nx_ * nx_;
        parallel_for_(leaves_.size(), copy_threads, [&](size_t l) {
            Vector<double>& u = nodes_[leaves_[l]]->data.vectorU();
            if ((size_t)u.size() != cells) u = Vector<double>(cells);     // repeated solves: the leaf's vector is reused
            std::memcpy(u.dataPointer(), &u_[l * cells], cells * sizeof(double));
        });
        if (!copy_back_operators) return;
        for (size_t i = 0; i < nodes_.size(); i++) fetch_vector_((int)i, EFGPU_VEC_G, nodes_[i]->data.vectorG());
    }

    void fetch_matrix_(int node, int which, Matrix<double>& M) {
        int r = 0, c = 0;
        check_(efgpu_operator_shape(h_, node, which, &r, &c), "efgpu_operator_shape");
        M = Matrix<double>(r, c);
        check_(efgpu_get_operator(h_, node, which, M.dataPointer(), (size_t)r * c), "efgpu_get_operator");
    }
    void fetch_vector_(int node, int which, Vector<double>& v) {
        int n = 0;
        check_(efgpu_vector_length(h_, node, which, &n), "efgpu_vector_length");
        v = Vector<double>(n);
        check_(efgpu_get_vector(h_, node, which, v.dataPointer(), (size_t)n), "efgpu_get_vector");
    }
};

}  // namespace EllipticForest

#endif  // ELLIPTIC_FOREST_B200_HPP_
