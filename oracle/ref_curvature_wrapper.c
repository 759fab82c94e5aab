/* ORACLE BUILD GLUE (test infrastructure).  The reference keeps c_curvature_grad `static`
 * (membrane_mesh_utils.c:915), so this translation unit textually includes the reference source
 * from where it lies (-I/root/reference/ch_shrinkwrap at build time; nothing is copied) and
 * re-exports it.  Output goes to oracle/_ref/ only. */
#include "membrane_mesh_utils.c"

void ref_curvature_grad(void *vertices, void *faces, void *halfedges, float dN, float skip_prob,
                        int n_vertices, float *k_0, float *k_1, float *e_0, float *e_1, float *H,
                        float *K, float *dH, float *dK, float *E, float *pE, float *dE_neighbors,
                        float kc, float kg, float c0, float *dEdN)
{
    c_curvature_grad(vertices, faces, (halfedge_t *)halfedges, dN, skip_prob, n_vertices, k_0, k_1,
                     e_0, e_1, H, K, dH, dK, E, pE, dE_neighbors, kc, kg, c0, (points_t *)dEdN);
}

/* c_holepunch_pair_candidate_faces is `static` too (membrane_mesh_utils.c:1301) */
void ref_holepunch_pair_candidate_faces(void *vertices, void *faces, void *halfedges, int *candidates,
                                        int n_candidates, int *pairs)
{
    c_holepunch_pair_candidate_faces(vertices, faces, (halfedge_t *)halfedges, candidates, n_candidates, pairs);
}
