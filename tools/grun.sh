#!/bin/bash
# gpurun from a frozen copy of the tree, with retries while the pod has no free slot (exit code 3: nothing charged).
#   tools/grun.sh <outfile> [gpurun options] -- '<command>'
# gpurun snapshots /root/repo when it gets a slot, which can be minutes after the call was made; so that edits made while a
# call waits cannot leak into it, the tree (sources + built .so) is copied to _stage/<id>/ first and the command runs there
# on the box, with gpurun_out/ linked back to the real one.  _stage/ is git-ignored.
set -u
out=$1; shift
opts=()
while [ "$1" != "--" ]; do opts+=("$1"); shift; done
shift
cmd="$*"
root=/root/repo
id=s$(date +%s)$$
mkdir -p $root/_stage/$id $root/gpurun_out
(cd $root && tar cf - --exclude=./.git --exclude=./gpurun_out --exclude=./_stage --exclude=./.pytest_cache --exclude='__pycache__' \
   --exclude='./multimodal_unlearning_b200/_build' --exclude='./*_r0*.json' . ) | (cd $root/_stage/$id && tar xf -)
wrapped="cd _stage/$id && rm -rf gpurun_out && ln -s \$GRAFT_REPO_ROOT/gpurun_out gpurun_out && $cmd"
rc=3
for i in $(seq 1 25); do
  /usr/local/graft/bin/gpurun "${opts[@]}" -- "$wrapped" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then break; fi
  if [ $rc -eq 2 ] && ! grep -q "another call" "$out"; then break; fi
  sleep 60
done
rm -rf $root/_stage/$id
exit $rc
