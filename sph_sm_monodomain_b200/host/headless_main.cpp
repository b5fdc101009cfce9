// headless_main.cpp — the reference's run protocol (main.cpp) without the GLUT window.
//
// Drives the drop-in SPH_SM_monodomain class through the exact call sequence main.cpp makes:
//   init()            main.cpp:490-496   new SPH_SM_monodomain(); init_cube() or init_mesh(file)
//   idle() x steps    main.cpp:318-352   turnOffStim() when half of the steps are left, then Animation()
//   display_points()  main.cpp:199-230   Get_Paticles() + Get_Particle_Number() every frame (here: every --frame-every steps)
//   exit_simulation() main.cpp:383-391   print_report(avg_fps, avg_step), delete
// so it doubles as the integration test of the class API and as the reference-facing end-to-end timing harness.
//
//   sphsm_headless [--cube | --csv FILE [--subsample 7] | --xyz FILE] [--steps 500] [--no-stim-off] [--quadratic]
//                  [--fused] [--frame-every K] [--dump FILE] [--capacity N --world X Y Z]
//
// --xyz reads raw little-endian float32 triples; --dump writes the final Particle[] array (132 bytes each).
#include <SPH_SM_monodomain.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

using std::vector;

// main.cpp:145-179: "x,y,z" per line; with subsample > 0 keep the first 3000 rows and then every subsample-th row
// (the reference switches this on by file name for biceps_simple_out_18475.csv and passes 7).
static bool read_cloud(const char *path, vector<m3Vector> &out, int subsample) {
    FILE *f = fopen(path, "r");
    if (!f) return false;
    float x, y, z;
    int got, row = 0;
    while ((got = fscanf(f, "%f,%f,%f\n", &x, &y, &z)) != EOF) {
        if (got == 3 && (subsample <= 0 || row < 3000 || row % subsample == 0)) out.push_back(m3Vector(x, y, z));
        row++;
    }
    fclose(f);
    return true;
}

static bool read_raw(const char *path, vector<m3Vector> &out) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    float v[3];
    while (fread(v, sizeof(float), 3, f) == 3) out.push_back(m3Vector(v[0], v[1], v[2]));
    fclose(f);
    return true;
}

// main.cpp:464-473: float loop variables advanced by the DOUBLE product kernel * 0.9; z outermost, x innermost
static void cube_positions(vector<m3Vector> &out) {
    const m3Vector W(1.5f, 1.5f, 1.5f);
    const float h = 0.04f;
    for (float k = W.z * 0.3f; k < W.z * 0.7f; k += h * 0.9)
        for (float j = W.y * 0.0f; j < W.y * 0.4f; j += h * 0.9)
            for (float i = W.x * 0.3f; i < W.x * 0.7f; i += h * 0.9) out.push_back(m3Vector(i, j, k));
}

int main(int argc, char **argv) {
    std::string csv, xyz, dump;
    bool cube = false, stim_off = true, quadratic = false, fused = false, staged = false;
    int steps = 500, subsample = 0, frame_every = 1, capacity = 0;
    float world[3] = {1.5f, 1.5f, 1.5f};
    for (int a = 1; a < argc; a++) {
        const std::string s = argv[a];
        if (s == "--cube") cube = true;
        else if (s == "--csv" && a + 1 < argc) csv = argv[++a];
        else if (s == "--xyz" && a + 1 < argc) xyz = argv[++a];
        else if (s == "--subsample" && a + 1 < argc) subsample = atoi(argv[++a]);
        else if (s == "--steps" && a + 1 < argc) steps = atoi(argv[++a]);
        else if (s == "--frame-every" && a + 1 < argc) frame_every = atoi(argv[++a]);
        else if (s == "--dump" && a + 1 < argc) dump = argv[++a];
        else if (s == "--capacity" && a + 1 < argc) capacity = atoi(argv[++a]);
        else if (s == "--world" && a + 3 < argc) { for (int k = 0; k < 3; k++) world[k] = (float)atof(argv[++a]); }
        else if (s == "--no-stim-off") stim_off = false;
        else if (s == "--quadratic") quadratic = true;
        else if (s == "--fused") fused = true;    // (the default since round 2; kept for old command lines)
        else if (s == "--staged") staged = true;  // the reference's seven separately timed stages
        else { fprintf(stderr, "unknown argument %s\n", s.c_str()); return 2; }
    }

    vector<m3Vector> positions;
    if (cube) cube_positions(positions);
    else if (!csv.empty()) { if (!read_cloud(csv.c_str(), positions, subsample)) { fprintf(stderr, "cannot open %s\n", csv.c_str()); return 2; } }
    else if (!xyz.empty()) { if (!read_raw(xyz.c_str(), positions)) { fprintf(stderr, "cannot open %s\n", xyz.c_str()); return 2; } }
    else cube = true, cube_positions(positions);

    SPH_SM_monodomain *sph = capacity > 0 ? new SPH_SM_monodomain(capacity, m3Vector(world[0], world[1], world[2])) : new SPH_SM_monodomain();
    sph->Init_Fluid(positions);
    if (cube) sph->turnOnStim_Cube(positions);  // init_cube, main.cpp:475-476
    else sph->turnOnStim_Mesh(positions);       // init_mesh, main.cpp:486-487
    if (quadratic) sph->flip_quadratic();
    if (fused) sph->set_stage_timing(false);
    if (staged) sph->set_stage_timing(true);

    duration_d stepping(0);
    double displacement_sum = 0.0;  // what display_points consumes; keeps the accessor honest
    int frames = 0;
    const auto run_t0 = std::chrono::system_clock::now();
    for (int left = steps; left > 0; left--) {
        const tpoint t0 = std::chrono::system_clock::now();
        if (stim_off && left == steps / 2) {  // main.cpp:329-334
            sph->turnOffStim();
            std::cout << "Turning stimulation off" << std::endl;
        }
        sph->Animation();
        sph->synchronize();  // the reference's step is synchronous: time the work, not the launch
        stepping += std::chrono::system_clock::now() - t0;
        if (frame_every > 0 && (steps - left) % frame_every == 0) {  // display(): Get_Paticles() + per-particle reads
            Particle *p = sph->Get_Paticles();
            const int n = sph->Get_Particle_Number();
            double s = 0.0;
            for (int i = 0; i < n; i++) s += p[i].getDisplacement();
            displacement_sum = n ? s / n : 0.0;
            frames++;
        }
    }
    sph->synchronize();
    const double wall = duration_d(std::chrono::system_clock::now() - run_t0).count();
    sph->print_report(wall > 0 ? frames / wall : 0.0, steps > 0 ? stepping.count() / steps : 0.0);
    fprintf(stderr, "headless: N=%d steps=%d wall=%.3fs mean displacement=%.6g\n", sph->Get_Particle_Number(), steps, wall, displacement_sum);

    if (!dump.empty()) {
        FILE *f = fopen(dump.c_str(), "wb");
        if (!f) { fprintf(stderr, "cannot write %s\n", dump.c_str()); return 2; }
        fwrite(sph->Get_Paticles(), sizeof(Particle), (size_t)sph->Get_Particle_Number(), f);
        fclose(f);
    }
    delete sph;
    return 0;
}
