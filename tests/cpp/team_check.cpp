// Host-logic check (no GPU): the tree decomposition the role / team / specialised kernels run on (csrc/tds_team.h) is
// built for a flat model read from a file and its invariants are verified.
//   team_check <model.bin> [n_act start_link]     exit 0 = consistent, 1 = violated (message on stderr), 4 = no decomposition
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tds_team.h"

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "team_check: " __VA_ARGS__); fprintf(stderr, "\n"); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::vector<double> m;
  {
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    double v;
    while (fread(&v, sizeof(double), 1, f) == 1) m.push_back(v);
    fclose(f);
  }
  DevModel D;
  int rc = tds_build_dev_model(m.data(), (int)m.size(), &D);
  if (rc) { fprintf(stderr, "team_check: model refused (%d)\n", rc); return 3; }
  tds_build_layout_w(&D, 4, 8, 4, -1);
  EnvParams E;
  memset(&E, 0, sizeof(E));
  if (argc >= 4) {
    E.n_act = atoi(argv[2]); E.start_link = atoi(argv[3]);
    int k = 0;
    for (int i = D.floating ? 0 : E.start_link; i < D.n_links && k < E.n_act; ++i)
      if (!(D.flags[i] & TDS_LF_FIXED)) E.act_link[k++] = i;
    E.n_act = k;
  }
  TeamModel TM;
  std::vector<TeamLink> tl;
  rc = tds_build_team(&D, &E, &TM, &tl);
  if (rc > 0) { printf("no decomposition (chain)\n"); return 4; }
  CHECK(rc == 0, "tds_build_team failed (%d)", rc);
  const int T = TDS_TEAM_T, n = D.n_links;
  auto at = [&](int r, int k) -> const TeamLink& { return tl[(size_t)r * TDS_TEAM_MAXK + k]; };
  // 1. every link is a trunk link (in every role's list, same position) or owned by exactly one role
  std::vector<int> owners(n, 0), trunk(n, 0);
  for (int r = 0; r < T; ++r) {
    CHECK(TM.n_loc[r] >= TM.n_trunk && TM.n_loc[r] <= TDS_TEAM_MAXK, "role %d: n_loc %d", r, TM.n_loc[r]);
    for (int k = 0; k < TM.n_loc[r]; ++k) {
      const TeamLink& L = at(r, k);
      CHECK(L.link >= 0 && L.link < n, "role %d pos %d: link %d", r, k, L.link);
      if (k < TM.n_trunk) { CHECK(L.link == at(0, k).link, "trunk lists differ at %d", k); trunk[L.link] = 1; }
      else ++owners[L.link];
      // 2. parents precede children in the role's list; adjacency flags are truthful
      const int p = D.parent[L.link];
      if (p < 0) CHECK(L.lpar == -1, "link %d: base parent but lpar %d", L.link, L.lpar);
      else { CHECK(L.lpar >= 0 && L.lpar < k && at(r, L.lpar).link == p, "link %d: lpar %d is not its parent %d", L.link, L.lpar, p); }
      if ((L.flags & TDS_TF_PARENT_ADJ) && p >= 0) CHECK(L.lpar == k - 1, "link %d: PARENT_ADJ but lpar %d != %d", L.link, L.lpar, k - 1);
      if (L.flags & TDS_TF_PARENT_TRUNK) CHECK(k >= TM.n_trunk && (p < 0 || trunk[p]), "link %d: PARENT_TRUNK on a non-attachment", L.link);
      CHECK(L.jtype == D.jtype[L.link] && L.q_idx == D.q_idx[L.link] && L.qd_idx == D.qd_idx[L.link], "link %d: copied fields differ", L.link);
      // 3. dof numbering: trunk dofs [0, n_td), own dofs [n_td, n_td + n_od), fixed joints none
      if (D.flags[L.link] & TDS_LF_FIXED) CHECK(L.ldof == -1, "fixed link %d has dof %d", L.link, L.ldof);
      else if (k < TM.n_trunk) CHECK(L.ldof >= (D.floating ? 6 : 0) && L.ldof < TM.n_td, "trunk link %d dof %d", L.link, L.ldof);
      else CHECK(L.ldof >= TM.n_td && L.ldof < TM.n_td + TM.n_od[r], "own link %d dof %d outside [%d, %d)", L.link, L.ldof, TM.n_td, TM.n_td + TM.n_od[r]);
      // 4. slots in range; a non-carried child has somewhere to add its inertia unless it hangs off a fixed base
      CHECK(L.acc_slot < TM.n_acc && L.par_slot < TM.n_acc, "link %d: slot out of range", L.link);
      if (!(L.flags & TDS_TF_PARENT_ADJ) && !(p < 0 && !D.floating)) CHECK(L.par_slot >= 0 || p < 0, "link %d: no accumulator for a non-adjacent child", L.link);
      if (L.par_slot >= 0 && L.lpar >= 0) CHECK(at(r, L.lpar).acc_slot == L.par_slot, "link %d: parent does not consume slot %d", L.link, L.par_slot);
    }
  }
  for (int i = 0; i < n; ++i) CHECK(trunk[i] ? owners[i] == 0 : owners[i] == 1, "link %d: trunk %d, owned %d times", i, trunk[i], owners[i]);
  // dofs of a role are distinct
  for (int r = 0; r < T; ++r) {
    std::vector<int> seen(TM.n_td + TM.n_od[r] + 1, 0);
    for (int k = 0; k < TM.n_loc[r]; ++k) { const int d = at(r, k).ldof; if (d >= 0) { CHECK(!seen[d], "role %d: dof %d twice", r, d); seen[d] = 1; } }
  }
  // 5. contact candidates: enumeration order of the reference (base geoms, then links in order), every candidate owned
  int cand = 0;
  for (int li = -1; li < n; ++li)
    for (int g = D.geom_begin[li + 1]; g < D.geom_begin[li + 2]; ++g) {
      const int pts = D.g_type[g] == TDSG_SPHERE ? 1 : (D.g_type[g] == TDSG_CAPSULE ? 2 : 0);
      for (int j = 0; j < pts && D.has_plane; ++j, ++cand) {
        const int r = TM.cand_owner[cand];
        CHECK(r >= 0 && r < T && TM.cand_lpt[cand] >= 0 && TM.cand_lpt[cand] < TM.n_pts[r], "candidate %d: owner %d lpt %d", cand, r, TM.cand_lpt[cand]);
        if (li >= 0 && !trunk[li]) {
          bool mine = false;
          for (int k = TM.n_trunk; k < TM.n_loc[r]; ++k) mine |= at(r, k).link == li;
          CHECK(mine, "candidate %d on link %d is not owned by the role of that link", cand, li);
        } else CHECK(r == 0, "candidate %d on the trunk / base must belong to role 0", cand);
      }
    }
  CHECK(cand == TM.n_cand, "candidate count %d != %d", cand, TM.n_cand);
  printf("ok: %d links, trunk %d, roles %d %d %d %d, trunk dofs %d, candidates %d\n", n, TM.n_trunk, TM.n_loc[0], TM.n_loc[1], TM.n_loc[2],
         TM.n_loc[3], TM.n_td, TM.n_cand);
  return 0;
}
