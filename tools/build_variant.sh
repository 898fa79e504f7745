#!/bin/bash
# Build a variant of the library next to the default one for A/B runs:
#   tools/build_variant.sh pflane -DMK_PF_LANE      -> markushgrapher_b200/lib/libmg_b200_pflane.so
#   tools/build_variant.sh fine   -DMK_FINE         -> ..._fine.so (in-kernel phase stamps, tools/mega_phase_profile.py)
# then on the GPU box:  MG_B200_LIB=$PWD/markushgrapher_b200/lib/libmg_b200_<name>.so python tools/ab_env.py ...
# The default library is rebuilt afterwards so the tree is left as it was.
set -e
NAME=$1; shift
cd "$(dirname "$0")/.."
MG_B200_CFLAGS="$*" python -c "from markushgrapher_b200 import build as b; b.build(force=True)"
cp markushgrapher_b200/lib/libmg_b200.so markushgrapher_b200/lib/libmg_b200_${NAME}.so
python -c "from markushgrapher_b200 import build as b; b.build(force=True)"
echo "built markushgrapher_b200/lib/libmg_b200_${NAME}.so"
