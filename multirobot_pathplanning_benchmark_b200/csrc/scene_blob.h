/* Scene blob word layout -- shared by the blob compiler
 * (multirobot_pathplanning_benchmark_b200/scene.py), the CUDA kernels and the fp64 CPU oracle
 * (oracle/oracle_scene.c).  Plain C.  A blob is an array of equally sized words: 4-byte
 * words (int32 / float) for the device, 8-byte words (int64 / double) for the oracle.
 *
 *   header  [MRB_HDR_WORDS]
 *   frames  [n_frames  * MRB_FRAME_WORDS]   { parent, joint code, q index, first shape | n shapes << 16, R[9] row-major, t[3] }
 *   shapes  [n_shapes  * MRB_SHAPE_WORDS]   { core, frame (-1 static), world offset, radius, data[16] }
 *           moving shapes first (sorted by frame), then static shapes (data in world coordinates)
 *           data: point c[3] | segment a[3] b[3] | box c[3] R[9] row-major half[3] | cylz c[3] r h
 *   chains  [n_chains * 2]                  { first frame, one past last frame } (serial joint paths)
 *           data[15] = radius of the bounding sphere around the centre (segment midpoint, box centre)
 *   pairs   8 typed lists of packed (a | b << 16 | kind << 28) shape indices, core(a) <= core(b);
 *           kind = broadphase bound: 0 bounding spheres, 1 / 2 face-normal bound against large box b
 *   static pairs [n * 3]                    { type, a, b }  both shapes static: constant per mode
 *   shape robot  [n_shapes]                 owning robot of a moving shape, -1 for static shapes
 */
#ifndef MRB_SCENE_BLOB_H
#define MRB_SCENE_BLOB_H

#define MRB_BLOB_MAGIC 0x4D524232
#define MRB_BLOB_VERSION 7
#define MRB_HDR_WORDS 136
#define MRB_FRAME_WORDS 16
#define MRB_SHAPE_WORDS 20
#define MRB_NUM_PAIR_TYPES 8

#define MRB_CORE_POINT 0
#define MRB_CORE_SEG 1
#define MRB_CORE_BOX 2
#define MRB_CORE_CYLZ 3 /* upright cylinder (axis = world z always): data c[3], r, half height */

/* pair types */
#define MRB_PT_POINT_POINT 0
#define MRB_PT_POINT_SEG 1
#define MRB_PT_SEG_SEG 2
#define MRB_PT_POINT_BOX 3
#define MRB_PT_SEG_BOX 4
#define MRB_PT_BOX_BOX 5
#define MRB_PT_CYLZ_CYLZ 6
#define MRB_PT_BOX_CYLZ 7 /* a = box (z-aligned, unrounded), b = cylinder */

/* joint codes */
#define MRB_J_HINGE_X 1
#define MRB_J_HINGE_Y 2
#define MRB_J_HINGE_Z 3
#define MRB_J_TRANS_XY_PHI 4
#define MRB_J_TRANS_X 5
#define MRB_J_TRANS_Y 6
#define MRB_J_TRANS_Z 7

/* header word indices */
#define MRB_H_MAGIC 0
#define MRB_H_VERSION 1
#define MRB_H_DOF 2
#define MRB_H_NFRAMES 3
#define MRB_H_NMOV 4
#define MRB_H_NSTA 5
#define MRB_H_WORLD_WORDS 6
#define MRB_H_NCHAINS 7
#define MRB_H_OFF_FRAMES 8
#define MRB_H_OFF_SHAPES 9
#define MRB_H_OFF_CHAINS 10
#define MRB_H_OFF_STATIC_PAIRS 11
#define MRB_H_N_STATIC_PAIRS 12
#define MRB_H_TOL 13        /* float */
#define MRB_H_STATIC_PEN 14 /* float: filled on the device by mrb200_scene_set_mode */
#define MRB_H_TOTAL_WORDS 15
#define MRB_H_NROBOTS 16
#define MRB_H_STAGED_WORDS 17 /* prefix the kernels copy into shared memory; the rest is read from global memory */
#define MRB_H_REC_BASE 18     /* first word of the broadphase records (staged) */
#define MRB_H_IDS_BASE 19     /* first word of the records' packed pair ids, one per record, in record order; inside the \
                                 staged prefix on small scenes, in the tail on large ones */
#define MRB_H_OFF_PAIRS 20 /* [20..27] */
#define MRB_H_N_PAIRS 28   /* [28..35] */
#define MRB_H_OFF_SHAPE_ROBOT 36
#define MRB_H_OFF_SCENTRE 37 /* static shape centres, 4 floats each */
#define MRB_H_IDS_STAGED 38  /* 1: the records' pair ids are part of the staged prefix */
#define MRB_H_GPTR 40        /* [40..41] device only: each CTA parks the blob's global address here after staging */
/* broadphase sublists [type][sublist] -> (record offset, count).  A record is 2 words
 * { (byte offset of moving shape X inside a W row) | (Y id << 16), float threshold }; record r (counted from
 * MRB_H_REC_BASE) has its packed (a | b << 16) pair ids for the narrowphase at word MRB_H_IDS_BASE + r.
 *   sublist 0: Y moving (id = its byte offset), bounding spheres, threshold = (bX + bY + slack)^2
 *   sublist 1: Y static (id = static index),    bounding spheres, same threshold
 *   sublist 2: Y a large static box (id = static index), separating-axis bound along the box's
 *              face normals; threshold = rX + rY + slack (X a segment) or bX + rY + slack */
#define MRB_H_BP 48
#define MRB_BP_SUBLISTS 3
#define MRB_H_BP_IDS 112 /* [type][sublist] -> first word of the sublist's packed pair ids (= MRB_H_IDS_BASE + record number) */

/* threshold below which a box-box edge-edge SAT axis (|a_i x b_j|^2) is skipped as degenerate */
#define MRB_SAT_PARALLEL_EPS2 1e-4

#endif
