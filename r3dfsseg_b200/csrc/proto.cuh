// internal launchers of proto.cu
#pragma once
#include "common.cuh"

// spill: fps_q8_spill_bytes(total rows of feat) bytes of scratch, or nullptr (streaming kernel)
int launch_fps_ex(const float* feat, int D, const int32_t* set_off, const int32_t* set_n,
                  int n_sets, int n_cap, int m_max, int k_for_count, int32_t* idx_out,
                  int32_t* cnt_out, cudaStream_t st, uint8_t* spill = nullptr,
                  int sets_per_group = 1, int impl = R3DFS_FPS_AUTO);
size_t fps_q8_spill_bytes(int64_t total_rows);
// cl = CTAs per set (0: the smallest cluster whose shared memory holds n_cap rows)
int launch_fps_q8(const float* feat, const int32_t* set_off, const int32_t* set_n, int n_launch,
                  int set_first, int set_per, int set_stride, int n_cap, int cl, int m_max,
                  int k_for_count, uint8_t* spill, int32_t* idx_out, int32_t* cnt_out,
                  cudaStream_t st);
int launch_multi_prototypes(const float* feat, int D, const int32_t* set_off,
                            const int32_t* set_n, int n_sets, int n_cap, int k, int32_t* picks,
                            int32_t* pick_cnt, int32_t* seeds, int32_t* proto_cnt,
                            int32_t* assign, float* partial, int32_t* pcount, float* seed_stats,
                            int sets_per_group, int64_t group_rows, float* proto_out, int ld_out,
                            cudaStream_t st, const StageRec* sr = nullptr,
                            uint8_t* fps_spill = nullptr, int fps_sets_per_group = 1);
// seed_stats: (2, n_sets, 128) floats of scratch
int launch_assign_tc(const float* feat, int D, const int32_t* set_off, const int32_t* set_n,
                     const int32_t* seeds, const int32_t* proto_cnt, int n_sets, int n_cap,
                     int m_max, int k, float* sn, float* ss, int32_t* assign, cudaStream_t st);
// scratch: partial (n_sets, chunks, k+1, D) floats and pcount (n_sets, chunks, k+1) ints with
// chunks = multi_prototypes_chunks(n_cap)
int multi_prototypes_chunks(int n_cap);
int launch_set_compaction(const float* F, int64_t ep_rows, int64_t sup_row_off, int E, int n_way,
                          int k_shot, int N, int D, const int32_t* sy, const int32_t* keep,
                          int32_t* fg_cnt, int32_t* set_off, int32_t* set_n, int32_t* cloud_bg_off,
                          int32_t* cloud_fg_off, float* setfeat, cudaStream_t st);
int launch_mdns(const float* sx, int64_t s_e, int64_t s_cloud, int64_t s_c, int64_t s_n,
                const int32_t* sy, const float* F, int64_t ep_rows, int64_t sup_row_off, int E,
                int n_way, int k_shot, int N, int D, float* cell_mean, int32_t* cell_cnt,
                int32_t* fg_cnt, int32_t* keep, float* clean_flag, cudaStream_t st,
                uint8_t* cell_mask_out = nullptr, float* degree_out = nullptr,
                float* scale_flag_out = nullptr);
