// CreateImageB200.cpp — validation + timing harness around the reference's own driver API.
//
//     CreateImageB200 [-iterations=N] [-methods=cpu,threads,b200,b200-multigpu] file.dat ...
//
// Does what the reference's CreateImage does (src/CreateImage.cpp:84-190: load the serialized
// create_image_struct, call RayTrace::create_image(info, method) per method, time it, run the
// reference's check_ans against the golden arrays embedded in the file, print the
// Avg/Min/Max/StdDev table) and ADDS the check BASELINE.json asks for: the image and I_ang of
// every non-CPU method are compared, two-sided and element-wise, with the result of the
// reference's own RayTraceImageCPU ("cpu" method) computed in the same process on the same input:
//     relative L2 error <= 1e-10 for both arrays, max element-wise relative error <= 1e-9 over
//     entries larger than 1e-6 of the array maximum.
// It links the UNMODIFIED reference sources plus the dispatcher with the two-branch registration
// of the B200 back-end (oracle/patch_dispatch.py).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "CreateImageHelpers.h"
#include "RayTrace.h"

static RayTrace::create_image_struct *load(const std::string &filename, double **image0, double **I_ang0)
{
    FILE *fid = fopen(filename.c_str(), "rb");
    if (!fid) {
        fprintf(stderr, "Error opening file: %s\n", filename.c_str());
        return NULL;
    }
    uint64_t n = 0;
    fread2(&n, sizeof(uint64_t), 1, fid);
    std::vector<char> data(n);
    fread2(data.data(), 1, n, fid);
    fclose(fid);
    RayTrace::create_image_struct *info = new RayTrace::create_image_struct();
    info->unpack(std::pair<const char *, size_t>(data.data(), n));
    *image0 = info->image;
    *I_ang0 = info->I_ang;
    info->image = NULL;
    info->I_ang = NULL;
    return info;
}

struct Errors {
    double rel_l2, max_rel;
};

static Errors compare(const double *a, const double *b, size_t n)
{
    double err = 0, norm = 0, amax = 0, worst = 0;
    for (size_t i = 0; i < n; i++) {
        err += (a[i] - b[i]) * (a[i] - b[i]);
        norm += b[i] * b[i];
        amax = std::max(amax, std::fabs(b[i]));
    }
    for (size_t i = 0; i < n; i++)
        if (std::fabs(b[i]) > 1e-6 * amax)
            worst = std::max(worst, std::fabs(a[i] - b[i]) / std::fabs(b[i]));
    Errors e = { norm > 0 ? std::sqrt(err / norm) : std::sqrt(err), worst };
    return e;
}

int main(int argc, char *argv[])
{
    Options options;
    std::vector<std::string> files = options.read_cmd(argc, argv);
    if (files.empty())
        return -2;
    std::vector<std::string> methods = options.methods;
    if (methods.empty()) {
        methods.push_back("cpu");
        methods.push_back("threads");
        methods.push_back("b200");
    }
    if (std::find(methods.begin(), methods.end(), "cpu") == methods.end())
        methods.insert(methods.begin(), "cpu"); // the oracle of the validation
    int N_errors = 0;
    for (size_t f = 0; f < files.size(); f++) {
        printf("\nRunning tests for %s\n\n", files[f].c_str());
        double *image0 = NULL, *I_ang0 = NULL;
        RayTrace::create_image_struct *info = load(files[f], &image0, &I_ang0);
        if (!info)
            return -2;
        const size_t n_img = (size_t) info->euv_beam->nx * info->euv_beam->ny * info->euv_beam->nv;
        const size_t n_ang = (size_t) info->euv_beam->na * info->euv_beam->nb;
        // warm-up of the GPU back-end (the reference does the same for its Cuda methods,
        // src/CreateImage.cpp:118-132)
        for (size_t m = 0; m < methods.size(); m++) {
            std::string lower = methods[m];
            std::transform(lower.begin(), lower.end(), lower.begin(), ::tolower);
            if (lower.substr(0, 4) == "b200" || lower.substr(0, 4) == "cuda") {
                RayTrace::create_image(info, methods[m]);
                free(info->image);
                free(info->I_ang);
                info->image = info->I_ang = NULL;
            }
        }
        std::vector<double> cpu_image, cpu_ang;
        std::vector<std::vector<double> > time(methods.size());
        for (size_t m = 0; m < methods.size(); m++) {
            printf("Running %s\n", methods[m].c_str());
            for (int it = 0; it < options.iterations; it++) {
                if (info->image)
                    free(info->image);
                if (info->I_ang)
                    free(info->I_ang);
                info->image = info->I_ang = NULL;
                const double start = getTime();
                RayTrace::create_image(info, methods[m]);
                time[m].push_back(getTime() - start);
            }
            if (image0 && I_ang0 && !check_ans(image0, I_ang0, *info))
                N_errors++;
            if (methods[m] == "cpu") {
                cpu_image.assign(info->image, info->image + n_img);
                cpu_ang.assign(info->I_ang, info->I_ang + n_ang);
            } else if (!cpu_image.empty()) {
                const Errors ei = compare(info->image, cpu_image.data(), n_img);
                const Errors ea = compare(info->I_ang, cpu_ang.data(), n_ang);
                const bool ok = ei.rel_l2 <= 1e-10 && ea.rel_l2 <= 1e-10 && ei.max_rel <= 1e-9 &&
                                ea.max_rel <= 1e-9;
                const bool gpu = methods[m].substr(0, 4) == "b200";
                printf("  vs RayTraceImageCPU: image relL2 %.3e max %.3e | I_ang relL2 %.3e max %.3e  %s\n",
                    ei.rel_l2, ei.max_rel, ea.rel_l2, ea.max_rel,
                    ok ? "PASS (<= 1e-10 / 1e-9)" : (gpu ? "FAIL" : "(informative)"));
                if (gpu && !ok)
                    N_errors++;
            }
        }
        printf("\n        METHOD    Avg     Min     Max   Std Dev\n");
        for (size_t m = 0; m < methods.size(); m++)
            printf("%14s %9.5f %9.5f %9.5f %9.5f\n", methods[m].c_str(), getAvg(time[m]),
                getMin(time[m]), getMax(time[m]), getDev(time[m]));
        free(image0);
        free(I_ang0);
        free(info->image);
        free(info->I_ang);
        info->image = info->I_ang = NULL;
        delete info->euv_beam;
        delete info->seed_beam;
        delete[] info->gain;
        delete info->seed;
        delete info;
    }
    printf(N_errors == 0 ? "\nAll tests passed\n" : "\nSome tests failed\n");
    return N_errors;
}
