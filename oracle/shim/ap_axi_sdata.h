// oracle/shim/ap_axi_sdata.h -- TEST INFRASTRUCTURE ONLY (oracle build).
//
// Stand-in for the Vivado-HLS AXI side-channel header included by
// streamtools.h:51.  Only qdma_axis<D,0,0,0> with get_data/set_data/set_keep/
// set_last is referenced (streamtools.h:1001-1037), and those adapters are out
// of scope -- the type exists so the header parses.
#ifndef FCB_ORACLE_SHIM_AP_AXI_SDATA_H
#define FCB_ORACLE_SHIM_AP_AXI_SDATA_H

#include "ap_int.h"

template <int D, int U, int TI, int TD> struct qdma_axis {
  ap_uint<D> data;
  ap_uint<(D + 7) / 8> keep;
  ap_uint<1> last;
  ap_uint<D> get_data() const { return data; }
  ap_uint<(D + 7) / 8> get_keep() const { return keep; }
  ap_uint<1> get_last() const { return last; }
  void set_data(const ap_uint<D>& d) { data = d; }
  void set_keep(const ap_uint<(D + 7) / 8>& k) { keep = k; }
  void set_last(const ap_uint<1>& l) { last = l; }
};

#endif
