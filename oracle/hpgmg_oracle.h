/*
 * hpgmg_oracle.h -- interface of the plain-C oracle (TEST INFRASTRUCTURE ONLY; see hpgmg_oracle.c).
 */
#ifndef HPGMG_ORACLE_H
#define HPGMG_ORACLE_H

#define OR_SHAPE_BOX        0
#define OR_SHAPE_STAR       1
#define OR_SHAPE_NO_CORNERS 2
#define OR_RESTRICT_CELL    0
#define OR_RESTRICT_FACE_I  1
#define OR_RESTRICT_FACE_J  2
#define OR_RESTRICT_FACE_K  3

typedef struct {
  int n, jS, kS, vol, nvec;   /* cells per side; strides and volume (doubles) of the padded box */
  double h, eig;              /* spacing; Gershgorin bound on lambda_max(D^-1 A)                */
  double **v;                 /* v[id] = padded [k][j][i] array                                 */
} olevel;

typedef struct {
  int nlevels, cheby, krylov;
  double a, b;
  olevel *L;
} ohier;

ohier *oracle_build(int log2_dim, int cheby);
void   oracle_destroy(ohier *H);
double oracle_fmg_solve(ohier *H, int onLevel, double *norm_of_F);
void   oracle_richardson(ohier *H, double norms[3], double *err, double *order);

void   oracle_apply_BCs_v2(const olevel *L, int id, int shape);
void   oracle_apply_BCs_v4(const olevel *L, int id, int shape);
void   oracle_extrapolate_betas(const olevel *L);
void   oracle_apply_op(const olevel *L, int Ax_id, int x_id, double b);
void   oracle_residual(const olevel *L, int res_id, int x_id, int rhs_id, double b);
void   oracle_smooth_gsrb(const olevel *L, int x_id, int rhs_id, double b);
void   oracle_smooth_cheby(const olevel *L, int x_id, int rhs_id, double b);
void   oracle_restriction(const olevel *Lc, int id_c, const olevel *Lf, int id_f, int type);
void   oracle_interpolation_v2(const olevel *Lf, int id_f, double prescale, const olevel *Lc, int id_c);
void   oracle_interpolation_v4(const olevel *Lf, int id_f, double prescale, const olevel *Lc, int id_c);
double oracle_norm(const olevel *L, int id);

double *oracle_vector(ohier *H, int level, int id);
int    oracle_level_dim(ohier *H, int level);
int    oracle_level_jstride(ohier *H, int level);
int    oracle_level_volume(ohier *H, int level);
double oracle_level_eig(ohier *H, int level);
int    oracle_num_levels(ohier *H);
int    oracle_krylov_iterations(ohier *H);
void   oracle_set_h(ohier *H, int level, double h);
olevel *oracle_level(ohier *H, int level);
#endif
