#!/bin/bash
# per kernel: number of tensor-core instructions (UTC*MMA) and of waterfall back-edges (BRA.U.ANY) in the SASS
cuobjdump -sass "${1:-tvqaplus_b200/libstage_b200.so}" | awk '/Function : /{name=$3} /UTC[A-Z]*MMA/{m[name]++} /BRA.U.ANY/{w[name]++} /UTCBAR|UTCCP/{c[name]++} END{for(n in m) printf "%-70s mma %3d  waterfall %3d\n", n, m[n], w[n]}' | c++filt | sed 's/stg:://g'
