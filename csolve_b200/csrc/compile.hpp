// compile.hpp -- host-side compiled model (owner of the arrays DevModel points to).
#pragma once
#include <string>
#include <vector>
#include "csolve_b200.h"
#include "device_model.h"

namespace csolve_dev {

struct CompiledModel {
  DevModel host;                       // pointers into the vectors below (host addresses)
  std::vector<ClauseRec> clause;
  std::vector<WatchRec> wrec;
  std::vector<int32_t> wrec_ptr;
  std::vector<unsigned long long> lov_pair;
  std::vector<int32_t> lov_cptr, lov_cval;
  std::vector<uint32_t> lov_fconst, lov_adj;
  std::vector<int32_t> watch_ptr, watch_idx, node_l, node_r, node_first, order, prio, root_dom;
  std::vector<uint8_t> node_op;
  std::vector<int32_t> sat_occ_ptr;
  std::vector<int2_t> sat_occ;
  std::vector<LinRel> linrel;
  std::vector<int2_t> dense_form;
  std::vector<LinClause> lin;
  std::vector<LinTerm> lin_term;
};

// returns CSOLVE_OK or an error code with a message in err
int compile_model(const csolve_flat_model &m, CompiledModel &out, std::string &err);

}  // namespace csolve_dev
