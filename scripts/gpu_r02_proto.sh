# tensor-core go/no-go prototypes. Usage: gpurun -- 'bash scripts/gpu_r02_proto.sh <tag>'
TAG=${1:-r02b}
mkdir -p gpurun_out
cd scripts/proto
timeout 120 ./mma_sync_rate > ../../gpurun_out/mma_sync_rate_$TAG.log 2>&1; cat ../../gpurun_out/mma_sync_rate_$TAG.log
timeout 120 ./layer_proto > ../../gpurun_out/layer_proto_$TAG.log 2>&1; cat ../../gpurun_out/layer_proto_$TAG.log
timeout 120 ./tc_layer_proto probe > ../../gpurun_out/tc_probe_$TAG.log 2>&1; echo "probe rc=$?"; cat ../../gpurun_out/tc_probe_$TAG.log
timeout 180 ./tc_layer_proto layer > ../../gpurun_out/tc_layer_$TAG.log 2>&1; echo "layer rc=$?"; cat ../../gpurun_out/tc_layer_$TAG.log
