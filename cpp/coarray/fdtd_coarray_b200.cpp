// fdtd_coarray_b200 -- the reference's distributed program (coarray/fdtd.F90) on B200s: one process ("image")
// per GPU, z-slab decomposition, halo planes over NCCL/NVLink instead of coarray remote reads.
//
//     ./fdtd_coarray_b200 [--images P] [Ni Nj Nk iterations]          (defaults: all visible GPUs, 512 512 512 25)
//
// Program structure follows coarray/fdtd.F90 line by line where it has an equivalent:
//   :30-33   this_image / num_images / next_img / pred_img     -> rank, nranks (ring neighbours live in the library)
//   :35-48   source period, active steps, source box            -> same expressions, 0-based
//   :51,149-164 init_decomposition                               -> fdtd_slab_range (remainder to the low images)
//   :70      init_fields                                         -> fdtd_create_ex zero-fills
//   :81-104  main loop: init_currents(t); update_B; exchange pred_Bx/By; update_E; exchange next_Ex/Ey; update_B
//                                                                -> fdtd_scatter (every image writes the cells it owns)
//                                                                   + fdtd_update_fields (exchanges inside, overlapped)
//   :106-109 J = 0                                               -> fdtd_zeroed_currents
//   :74-78,130-135 timing on image 1, "Running on N images", "Total execution time"
//   :137-139,292-302 print_full_E_slice: 10x10 Ex values at k = Nk/2 read from the owning image
// No MPI / Fortran runtime: the images are forked from one launcher, the NCCL unique id and the printed slice travel
// through an anonymous shared mapping.  (J sign: the Fortran uses `- coef_J*J` with coef_J = +4*PI*dt, the C++
// classes `+ cJ*J` with cJ = -4*PI*dt; same value.)
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fdtd_b200.h"

namespace {

const double C = 3e10, PI = 3.14159265358;   // coarray/fdtd.F90:8 (= include/Constants.h:6,11)

struct Shared {
    std::atomic<int> id_ready;
    std::atomic<int> arrived[4];      // sense-free counting barriers, one per use
    std::atomic<int> failed;
    char nccl_id[FDTD_NCCL_UNIQUE_ID_BYTES];
    double slice[100];
};

void barrier(Shared* sh, int which, int n) {
    sh->arrived[which].fetch_add(1);
    while (sh->arrived[which].load() < n && !sh->failed.load()) usleep(50);
}

#define CHECK(call)                                                                       \
    do {                                                                                  \
        if ((call) != FDTD_OK) {                                                          \
            std::fprintf(stderr, "image %d: %s failed: %s\n", img, #call, fdtd_last_error()); \
            sh->failed.store(1);                                                          \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

int run_image(Shared* sh, int img, int nimg, int Ni, int Nj, int Nk, int num_iterations) {
    const double dx = C, dy = C, dz = C, dt = 0.2;
    const double TT = 8.0, Tx = 4 * C, Ty = 4 * C, Tz = 4 * C;
    const int current_time = std::min(static_cast<int>(TT / dt), num_iterations);   // :39
    const double bnd_i = Ni / 2.0 * dx, bnd_j = Nj / 2.0 * dy, bnd_k = Nk / 2.0 * dz;
    // :43-49 (1-based, inclusive) -> 0-based, exclusive upper bound: the same cells as sample.cpp:57-63
    const int start_i = static_cast<int>(std::floor((-Tx / 4.0 + bnd_i) / dx)), max_i = static_cast<int>(std::floor((Tx / 4.0 + bnd_i) / dx));
    const int start_j = static_cast<int>(std::floor((-Ty / 4.0 + bnd_j) / dy)), max_j = static_cast<int>(std::floor((Ty / 4.0 + bnd_j) / dy));
    const int start_k = static_cast<int>(std::floor((-Tz / 4.0 + bnd_k) / dz)), max_k = static_cast<int>(std::floor((Tz / 4.0 + bnd_k) / dz));

    fdtd_config_t cfg;
    fdtd_config_init(&cfg);
    cfg.grid = fdtd_params_t{Ni, Nj, Nk, -bnd_i, bnd_i, -bnd_j, bnd_j, -bnd_k, bnd_k, dx, dy, dz};
    cfg.dt = dt;
    cfg.device = img;
    cfg.rank = img;
    cfg.nranks = nimg;
    fdtd_solver_t* s = nullptr;
    CHECK(fdtd_create_ex(&cfg, &s));
    if (nimg > 1) {
        if (img == 0) {
            CHECK(fdtd_nccl_unique_id(sh->nccl_id, sizeof(sh->nccl_id)));
            sh->id_ready.store(1);
        }
        while (!sh->id_ready.load() && !sh->failed.load()) usleep(50);
        if (sh->failed.load()) return 1;
        CHECK(fdtd_comm_init(s, sh->nccl_id, sizeof(sh->nccl_id)));
    }

    std::vector<int64_t> idx;
    for (int k = start_k; k < max_k; ++k)
        for (int j = start_j; j < max_j; ++j)
            for (int i = start_i; i < max_i; ++i) idx.push_back(i + (int64_t)j * Ni + (int64_t)k * Ni * Nj);
    std::vector<double> val(idx.size());

    CHECK(fdtd_sync(s));
    barrier(sh, 0, nimg);
    if (img == 0) { std::printf("Running on %d images\n", nimg); std::fflush(stdout); }
    const auto t0 = std::chrono::steady_clock::now();

    for (int t = 1; t <= current_time; ++t) {
        // init_currents(t), :182-204 -- value = sin(2 PI t dt / TT) cos^2(2 PI i dx / Tx) cos^2(...) cos^2(...)
        size_t q = 0;
        for (int k = start_k; k < max_k; ++k)
            for (int j = start_j; j < max_j; ++j)
                for (int i = start_i; i < max_i; ++i)
                    val[q++] = ((std::sin(2.0 * PI * (t * dt) / TT) * std::pow(std::cos(2.0 * PI * (i * dx) / Tx), 2.0)) *
                                std::pow(std::cos(2.0 * PI * (j * dy) / Ty), 2.0)) * std::pow(std::cos(2.0 * PI * (k * dz) / Tz), 2.0);
        for (int c = FDTD_JX; c <= FDTD_JZ; ++c) CHECK(fdtd_scatter(s, c, idx.data(), val.data(), idx.size()));
        CHECK(fdtd_update_fields(s));
    }
    CHECK(fdtd_zeroed_currents(s));                                   // :106-109
    if (num_iterations > current_time) CHECK(fdtd_step(s, num_iterations - current_time));   // :111-128
    CHECK(fdtd_sync(s));
    barrier(sh, 1, nimg);                                             // sync all, :130
    if (img == 0) {
        const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("Total execution time: %.2f seconds\n", el);
        std::fflush(stdout);
    }

    // print_full_E_slice, :292-302: rows j = Nj/2-5 .. Nj/2+4, columns i = Ni/2-5 .. Ni/2+4 at plane Nk/2 (0-based)
    fdtd_info_t info;
    CHECK(fdtd_get_info(s, &info));
    const int kk = Nk / 2;
    if (kk >= info.k_begin && kk < info.k_end) {
        std::vector<int64_t> pidx;
        for (int j = Nj / 2 - 5; j < Nj / 2 + 5; ++j)
            for (int i = Ni / 2 - 5; i < Ni / 2 + 5; ++i) pidx.push_back(i + (int64_t)j * Ni + (int64_t)kk * Ni * Nj);
        CHECK(fdtd_gather(s, FDTD_EX, pidx.data(), sh->slice, pidx.size()));
    }
    barrier(sh, 2, nimg);
    if (img == 0) {
        for (int r = 0; r < 10; ++r) {
            for (int c = 0; c < 10; ++c) std::printf("%12.5f", sh->slice[r * 10 + c]);
            std::printf("\n");
        }
        std::fflush(stdout);
    }
    barrier(sh, 3, nimg);
    fdtd_destroy(s);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    int images = 0, dims[4] = {512, 512, 512, 25}, nd = 0;   // coarray/fdtd.F90:6-7
    for (int a = 1; a < argc; ++a) {
        if (!std::strcmp(argv[a], "--images") && a + 1 < argc) images = std::atoi(argv[++a]);
        else if (nd < 4) dims[nd++] = std::atoi(argv[a]);
    }
    if (nd != 0 && nd != 4) {
        std::fprintf(stderr, "usage: %s [--images P] [Ni Nj Nk iterations]\n", argv[0]);
        return 1;
    }
    if (images <= 0) {
        // count GPUs in a throw-away child so that the launcher itself never creates a CUDA context before fork()
        int fds[2];
        if (pipe(fds) != 0) return 1;
        pid_t pid = fork();
        if (pid == 0) {
            int n = 0;
            fdtd_config_t cfg;
            fdtd_config_init(&cfg);
            cfg.grid = fdtd_params_t{4, 4, 4, 0, 1, 0, 1, 0, 1, 1, 1, 1};
            cfg.dt = 0.1;
            for (; n < 64; ++n) {
                cfg.device = n;
                fdtd_solver_t* s = nullptr;
                if (fdtd_create_ex(&cfg, &s) != FDTD_OK) break;
                fdtd_destroy(s);
            }
            if (write(fds[1], &n, sizeof(n)) != sizeof(n)) _exit(1);
            _exit(0);
        }
        int n = 0;
        if (read(fds[0], &n, sizeof(n)) != sizeof(n)) n = 0;
        waitpid(pid, nullptr, 0);
        images = n;
        if (images <= 0) {
            std::fprintf(stderr, "no CUDA device available: fdtd_coarray_b200 has no CPU fallback\n");
            return 1;
        }
    }
    void* mem = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (mem == MAP_FAILED) return 1;
    Shared* sh = new (mem) Shared();
    std::memset(sh->slice, 0, sizeof(sh->slice));
    std::vector<pid_t> kids;
    for (int img = 0; img < images; ++img) {
        pid_t pid = fork();
        if (pid == 0) _exit(run_image(sh, img, images, dims[0], dims[1], dims[2], dims[3]));
        kids.push_back(pid);
    }
    int rc = 0;
    for (pid_t pid : kids) {
        int st = 0;
        waitpid(pid, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 1;
    }
    return rc;
}
