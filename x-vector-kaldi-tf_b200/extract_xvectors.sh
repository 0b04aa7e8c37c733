#!/bin/bash
# extract_xvectors.sh <nnet-dir> <data-dir> <xvector-dir>
#
# GPU counterpart of the reference's local/tf/extract_xvectors.sh: same positional arguments, same option names where they
# still mean something, same stages and the same files left behind (xvector.<job>.{ark,scp}, xvector.scp,
# spk_xvector.{ark,scp}, num_utts.ark, log/extract.<job>.log), so sid-style recipes can call it unchanged.
#
# The reference splits the data directory into --nj pieces and starts one TensorFlow process per piece
# (local/tf/extract_xvectors.sh:63-88).  Here a job is one process of this directory's extract_embedding.py bound to one
# GPU (job j -> device (j-1) mod num-gpus through XVEC_DEVICE); one process saturates a B200, so --nj defaults to the
# number of visible GPUs.  Features come through the same Kaldi pipe the reference builds (sliding-window CMN over 300
# frames, then voiced frames only; local/tf/extract_xvectors.sh:68).

set -o pipefail

# ---- options (Kaldi's parse_options.sh turns --foo-bar X into foo_bar=X) ----
num_gpus=$(nvidia-smi -L 2>/dev/null | grep -c '^GPU')
if [ -z "${num_gpus}" ] || [ "${num_gpus}" -lt 1 ]; then num_gpus=1; fi
nj=                 # jobs; empty = one per GPU
cmd=run.pl
chunk_size=-1       # <= 0: take max_chunk_size of the nnet dir
use_gpu=true        # accepted for compatibility; this build has no CPU path
device_frontend=true  # true: sliding CMVN + voiced-frame selection run on the GPU (raw feats.scp + vad.scp go in);
                      # false: the reference's Kaldi pipe (apply-cmvn-sliding | select-voiced-frames) feeds the job
stage=0

echo "$0 $*"

script_dir=$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)
[ -f path.sh ] && . ./path.sh
. parse_options.sh || exit 1

die() { echo "$0: $*" >&2; exit 1; }

if [ $# -ne 3 ]; then
  cat >&2 <<USAGE
Usage: $0 [options] <nnet-dir> <data> <xvector-dir>
 e.g.: $0 exp/xvector_nnet data/train exp/xvectors_train
Options:
  --cmd (utils/run.pl|utils/queue.pl <queue opts>)   how to run the jobs
  --num-gpus <n>      GPUs to use (default: all visible)
  --nj <n>            number of jobs, one process each (default: --num-gpus)
  --stage <n>         0 extract, 1 merge scp files, 2 speaker means
  --chunk-size <n>    frames per chunk; chunks of an utterance are averaged (default: max_chunk_size of the nnet dir)
  --device-frontend (true|false)   CMVN + VAD selection on the GPU instead of the Kaldi pipe (default: true)
USAGE
  exit 1
fi

nnet_dir=$1
data_dir=$2
out_dir=$3
[ -n "${nj}" ] || nj=${num_gpus}

for required in "${nnet_dir}/model_final/model.meta" "${nnet_dir}/min_chunk_size" "${nnet_dir}/max_chunk_size" \
                "${data_dir}/feats.scp" "${data_dir}/vad.scp"; do
  [ -f "${required}" ] || die "No such file ${required}"
done

min_chunk=$(cat "${nnet_dir}/min_chunk_size")
max_chunk=$(cat "${nnet_dir}/max_chunk_size")
[ "${chunk_size}" -gt 0 ] || chunk_size=${max_chunk}
[ "${chunk_size}" -le "${max_chunk}" ] || die "specified chunk size of ${chunk_size} is larger than the maximum chunk size, ${max_chunk}"

mkdir -p "${out_dir}/log"
utils/split_data.sh --per-utt "${data_dir}" "${nj}" || exit 1
split_dir=${data_dir}/split${nj}utt

# feature pipe of job $1 (what extract_embedding.py opens as its --feature-rspecifier)
feature_pipe() {
  local part=${split_dir}/$1
  echo "apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300 scp:${part}/feats.scp ark:- |" \
       "select-voiced-frames ark:- scp,s,cs:${part}/vad.scp ark:- |"
}

if [ "${stage}" -le 0 ]; then
  echo "$0: extracting xvectors for ${data_dir}: ${nj} job(s) on ${num_gpus} GPU(s)"
  pids=()
  for job in $(seq "${nj}"); do
    if [ "${device_frontend}" = true ]; then
      # same options the pipe passes to apply-cmvn-sliding; select-voiced-frames becomes --vad-rspecifier
      feature_args=(--feature-rspecifier="scp:${split_dir}/${job}/feats.scp" --apply-cmvn-sliding=yes --cmn-window=300
                    --norm-vars=false --center=true --vad-rspecifier="scp,s,cs:${split_dir}/${job}/vad.scp")
    else
      feature_args=(--feature-rspecifier="$(feature_pipe "${job}")")
    fi
    XVEC_DEVICE=$(( (job - 1) % num_gpus )) ${cmd} "${out_dir}/log/extract.${job}.log" \
      python "${script_dir}/extract_embedding.py" --use-gpu=yes \
        --min-chunk-size="${min_chunk}" --chunk-size="${chunk_size}" "${feature_args[@]}" \
        --vector-wspecifier="| copy-vector ark:- ark,scp:${out_dir}/xvector.${job}.ark,${out_dir}/xvector.${job}.scp" \
        --model-dir="${nnet_dir}/model_final" &
    pids+=($!)
  done
  failed=0
  for pid in "${pids[@]}"; do wait "${pid}" || failed=$((failed + 1)); done
  [ "${failed}" -eq 0 ] || die "${failed} extraction job(s) failed; see ${out_dir}/log/extract.*.log"
fi

if [ "${stage}" -le 1 ]; then
  echo "$0: combining xvectors across jobs"
  : > "${out_dir}/xvector.scp"
  for job in $(seq "${nj}"); do
    cat "${out_dir}/xvector.${job}.scp" >> "${out_dir}/xvector.scp" || exit 1
  done
fi

if [ "${stage}" -le 2 ]; then
  echo "$0: computing mean of xvectors for each speaker"
  run.pl "${out_dir}/log/speaker_mean.log" \
    ivector-mean "ark:${data_dir}/spk2utt" "scp:${out_dir}/xvector.scp" \
      "ark,scp:${out_dir}/spk_xvector.ark,${out_dir}/spk_xvector.scp" "ark,t:${out_dir}/num_utts.ark" || exit 1
fi
