/* The drop-in boundary used from plain C: no Python, no torch, only include/des_b200.h and libdes_b200.so.
 *
 *   gcc -std=c99 -Iinclude examples/nes_session.c -Ldistributedes_b200 -ldes_b200 -Wl,-rpath,$PWD/distributedes_b200 -o nes_session
 *   ./nes_session [generations] [population] [hidden] [precision 0|1|2]
 *
 * Runs `generations` NES generations (the loop body of natural_es.py:62-96) of a 2-hidden-layer tanh MLP on a
 * deterministic synthetic observation tape with host buffers, printing the mean fitness per generation and a checksum of
 * the final parameters.  Exit codes: 0 ok, 3 no CUDA device (the library has no CPU fallback), 1 any other error. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "des_b200.h"

static uint32_t lcg(uint32_t *s) { *s = *s * 1664525u + 1013904223u; return *s; }
static float unif(uint32_t *s) { return ((lcg(s) >> 8) + 0.5f) * (1.0f / 16777216.0f) * 2.0f - 1.0f; }   /* (-1, 1) */

int main(int argc, char **argv) {
    const int gens = argc > 1 ? atoi(argv[1]) : 3;
    const int64_t N = argc > 2 ? atoll(argv[2]) : 256;
    const int H = argc > 3 ? atoi(argv[3]) : 64;
    const int precision = argc > 4 ? atoi(argv[4]) : DES_FWD_FP32;
    const des_dims dims = {24, H, 4, 128};
    const int64_t P = des_param_count(dims.state_dim, dims.hidden, dims.action_dim);
    if (P <= 0) { fprintf(stderr, "bad dims\n"); return 1; }
    float *theta = malloc(sizeof(float) * P), *obs = malloc(sizeof(float) * dims.tape_len * dims.state_dim);
    float *target = malloc(sizeof(float) * dims.tape_len * dims.action_dim), *fitness = malloc(sizeof(float) * N);
    uint32_t s = 12345u;
    for (int64_t j = 0; j < P; ++j) theta[j] = 0.1f * unif(&s);
    for (int i = 0; i < dims.tape_len * dims.state_dim; ++i) obs[i] = 1.5f * unif(&s);
    for (int i = 0; i < dims.tape_len * dims.action_dim; ++i) target[i] = 0.9f * unif(&s);

    const des_opt opt = {0.1, 0.1, 0.005, 0.9, 0.999, 1e-8};      /* sigma, lr (natural_es.py:143-144), wd, Adam (utils.py:150-157) */
    des_session *sess = NULL;
    int rc = des_session_create(&sess, 0, dims, N, 0, N, opt, 1.0, 7, precision, theta);
    if (rc != DES_OK) {
        fprintf(stderr, "des_session_create: %d: %s\n", rc, des_last_error());
        return rc == DES_ERR_NO_DEVICE ? 3 : 1;
    }
    for (int g = 0; g < gens; ++g) {
        rc = des_session_generation_host(sess, obs, target, NULL, fitness, NULL, theta);
        if (rc != DES_OK) { fprintf(stderr, "generation %d: %d: %s\n", g, rc, des_last_error()); des_session_destroy(sess); return 1; }
        double mean = 0.0;
        for (int64_t i = 0; i < N; ++i) mean += fitness[i];
        printf("generation %d mean_fitness %.6f\n", g, mean / (double)N);
    }
    double chk = 0.0;
    for (int64_t j = 0; j < P; ++j) chk += (double)theta[j] * (double)((j % 7) + 1);
    printf("theta_checksum %.9e\n", chk);
    des_session_destroy(sess);
    free(theta); free(obs); free(target); free(fitness);
    return 0;
}
