/*
 * sgtd_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a dependency-free CPU restatement of the
 * reference's (Hfx-J/SGTD) one-shot localization hot path.  It exists to check
 * the CUDA product (sgtd_b200/) and to be timed as the `cpu_baseline` /
 * `--impl reference` arm of bench.py.  Nothing under sgtd_b200/ may include,
 * link or call it.
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN CODE, with one stated gap.  The
 * reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
 * section 4 / 8c) and its build needs ROS, PCL, FLANN, Eigen and Ceres, none of
 * which is installed.  Its hot-path sources themselves (src/STDesc.cpp,
 * include/desc/STDesc.h, include/cluster_manager.hpp) do compile unmodified
 * against small stand-in headers for those libraries (oracle/shim/): that build is
 * oracle/_ref/libsgtd_ref.so (oracle/Makefile, oracle/ref_wrap.cpp), and
 * tests/test_reference_build.py requires this restatement to equal it -- integer
 * outputs equal, descriptors and poses byte-equal -- on seeded worlds, degenerate
 * worlds and alternative configs; tests/golden/ref_*.npz are vectors written by
 * that build.  The gap: third-party ARITHMETIC is the shim's, restated from the
 * published algorithms (FLANN exact kNN with L2_Simple<float>, Eigen JacobiSVD,
 * Eigen's fixed-size reduction order e0 + (e1 + e2)); gen_labels / gen_graphs
 * (src/get_json.cpp) drag in the whole node and stay restated, anchored on source
 * lines.  Each function cites the reference file:line it follows.
 *
 * Reference paths are relative to /root/reference/src/sgtd/ :
 *   R/src/STDesc.cpp, R/include/desc/STDesc.h, R/include/cluster_manager.hpp,
 *   R/src/get_json.cpp
 */
#ifndef SGTD_ORACLE_H
#define SGTD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The 7 live ConfigSetting fields (R/include/desc/STDesc.h:38-72). */
typedef struct orc_config {
  int32_t descriptor_near_num;   /* 10  */
  int32_t candidate_num;         /* 50  */
  double descriptor_min_len;     /* 0.5 */
  double descriptor_max_len;     /* 50  */
  double std_side_resolution;    /* 1   */
  double rough_dis_threshold;    /* 0.03*/
  double icp_threshold;          /* 0.4 */
} orc_config;

/* Compact restatement of STDesc (R/include/desc/STDesc.h:75-97); 72 bytes. */
typedef struct orc_desc {
  double side[3];   /* side_length_ = scale * sorted sides                 */
  float vert[9];    /* vertex_A_, vertex_B_, vertex_C_ (exact float values) */
  uint32_t frame;   /* frame_id_                                           */
  uint8_t lab[3];   /* vertex_attached_ (labels of A,B,C)                  */
  uint8_t pad;
  uint16_t anchor;  /* node_id[0] = i                                      */
  uint8_t m, n;     /* node_id[1..2] = kNN ranks (sic, reference bug)      */
} orc_desc;

/* One verified candidate == one LOOP_RESULT (R/include/desc/STDesc.h:99-104)
 * plus the selector's bookkeeping. */
typedef struct orc_cand {
  int32_t frame;      /* match_id_.second                         */
  int32_t votes;      /* match_array[frame] when selected         */
  int32_t nmatch;     /* match_list_.size()                       */
  int32_t score;      /* (int)verify_score : #inliers or -1       */
  int32_t match_off;  /* offset into the match arrays             */
  int32_t inlier_off; /* offset into the inlier array             */
  int32_t ninlier;
  int32_t best_hyp;   /* max_vote_index (hypothesis), -1 if none  */
  double R[9];        /* row-major rotation                        */
  double t[3];
} orc_cand;

typedef struct orc_search_stats {
  int64_t Q; /* query descriptors                                            */
  int64_t P; /* probes passing the 1.5-ball test (incl. duplicate cells)     */
  int64_t Pfound; /* ... that found a bucket                                 */
  int64_t E; /* DB entries scanned                                           */
  int64_t M; /* matches emitted                                              */
} orc_search_stats;

typedef struct orc_handle orc_handle;

orc_handle *orc_create(const orc_config *cfg);
void orc_destroy(orc_handle *h);
uint32_t orc_current_frame_id(const orc_handle *h);
int64_t orc_db_size(const orc_handle *h);

/* BuildSingleScanSTD (R/src/STDesc.cpp:174-315).  xyz = K x 3 floats, label =
 * K uint32.  Writes up to cap descriptors, returns the number produced (may
 * exceed cap: call again) or -1 if K < descriptor_near_num. */
int64_t orc_build(orc_handle *h, const float *xyz, const uint32_t *label,
                  int32_t K, orc_desc *out, int64_t cap);

/* AddSTDescs (R/src/STDesc.cpp:149-172). */
void orc_add(orc_handle *h, const orc_desc *d, int64_t n);

/* Build + add of many keyframes at once (scan s = nodes off[s]..off[s+1]); builds run on nthreads
 * threads, adds in scan order.  Same database as nscans orc_build / orc_add calls. */
int64_t orc_build_add_many(orc_handle *h, const float *xyz, const uint32_t *label, const int64_t *off,
                           int32_t nscans, int32_t nthreads);

/* DB key of a descriptor as AddSTDescs forms it; returns x,y,z,code. */
void orc_db_key(const orc_desc *d, int32_t out[4]);

/* SearchLoop (R/src/STDesc.cpp:84-147) = candidate_selector (:318-460) +
 * candidate_verify (:462-547) + triangle_solver (:549-571).
 *   cands[cap_cand]        one per candidate, in selector order
 *   m_q, m_cell, m_g[cap_match]   match lists (query desc idx, probe ordinal
 *                           0..26, global DB insertion index), concatenated
 *   inl[cap_match]         inlier indices (into the candidate's match list)
 *   votes_out[n_frames]    optional (may be NULL): match_array before ranking
 *   best[2]                loop_result: (frame or -1, score)
 * returns the number of candidates, or -1 on empty input, -2 on capacity.    */
int32_t orc_search(orc_handle *h, const orc_desc *q, int64_t nq,
                   orc_cand *cands, int32_t cap_cand, int32_t *m_q,
                   uint8_t *m_cell, uint32_t *m_g, int32_t *inl,
                   int64_t cap_match, int32_t *votes_out, int64_t n_frames,
                   double best[2], orc_search_stats *stats, int32_t nthreads);

/* triangle_solver on one pair (for unit tests). */
void orc_triangle_solver(const orc_desc *src, const orc_desc *ref, double R[9],
                         double t[3]);

/* 3x3 SVD restating Eigen::JacobiSVD (for unit tests): A = U diag(s) V^T,
 * row-major. */
void orc_jacobi_svd3(const double A[9], double U[9], double s[3], double V[9]);

#ifdef __cplusplus
}
#endif
#endif
