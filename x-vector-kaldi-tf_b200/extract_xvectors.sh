#!/bin/bash
# GPU counterpart of the reference's local/tf/extract_xvectors.sh (same positional arguments, options, stages and output
# files: xvector.JOB.{ark,scp}, xvector.scp, spk_xvector.{ark,scp}, num_utts.ark).
#
# What differs from the reference (local/tf/extract_xvectors.sh:63-95): there the data directory is split into --nj pieces
# and every piece is a separate TensorFlow process.  Here --nj is the number of GPU processes: each job reads its split
# through the same Kaldi feature pipe (apply-cmvn-sliding | select-voiced-frames) and runs extract_embedding.py from this
# directory on GPU (job-1) mod --num-gpus; one process keeps a whole GPU busy, so --nj defaults to the number of GPUs.
# (One multi-GPU process over ONE pipe is the other way to run it: torchrun --nproc-per-node N extract_embedding.py ...)

# Begin configuration section.
num_gpus=$(nvidia-smi -L 2>/dev/null | wc -l)
[ "${num_gpus}" -ge 1 ] 2>/dev/null || num_gpus=1
nj=${num_gpus}
cmd="run.pl"

chunk_size=-1     # The chunk size over which the embedding is extracted.
                  # If left unspecified, it uses the max_chunk_size in the nnet directory.
use_gpu=true      # kept for command-line compatibility: there is no CPU path
stage=0

echo "${0} $@"  # Print the command line for logging

here=$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)

if [ -f path.sh ]; then . ./path.sh; fi
. parse_options.sh || exit 1;

if [ $# != 3 ]; then
  echo "Usage: ${0} <nnet-dir> <data> <xvector-dir>"
  echo " e.g.: ${0} exp/xvector_nnet data/train exp/xvectors_train"
  echo "main options (for others, see top of script file)"
  echo "  --cmd (utils/run.pl|utils/queue.pl <queue opts>) # how to run jobs."
  echo "  --num-gpus <n|all visible>                       # GPUs to spread the jobs over"
  echo "  --nj <n|num-gpus>                                # Number of jobs (one process per job)"
  echo "  --stage <stage|0>                                # To control partial reruns"
  echo "  --chunk-size <n|-1>                              # If provided, extracts embeddings with specified"
  echo "                                                   # chunk size, and averages to produce final embedding"
  exit 1;
fi

srcdir=$1
data=$2
dir=$3

for f in ${srcdir}/model_final/model.meta ${srcdir}/min_chunk_size ${srcdir}/max_chunk_size ${data}/feats.scp ${data}/vad.scp ; do
  [ ! -f ${f} ] && echo "No such file $f" && exit 1;
done

min_chunk_size=`cat ${srcdir}/min_chunk_size 2>/dev/null`
max_chunk_size=`cat ${srcdir}/max_chunk_size 2>/dev/null`

model_dir=${srcdir}/model_final

if [ ${chunk_size} -le 0 ]; then
  chunk_size=${max_chunk_size}
fi

if [ ${max_chunk_size} -lt ${chunk_size} ]; then
  echo "${0}: specified chunk size of ${chunk_size} is larger than the maximum chunk size, ${max_chunk_size}" && exit 1;
fi

mkdir -p ${dir}/log

utils/split_data.sh --per-utt ${data} ${nj}
echo "${0}: extracting xvectors for ${data}"
sdata=${data}/split${nj}utt/JOB

# Set up the features
feature_rspecifier="apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300 scp:${sdata}/feats.scp ark:- | select-voiced-frames ark:- scp,s,cs:${sdata}/vad.scp ark:- |"

if [ ${stage} -le 0 ]; then
  echo "${0}: extracting xvectors from nnet on ${num_gpus} GPU(s), ${nj} job(s)"
  for g in $(seq ${nj}); do
    XVEC_DEVICE=$(( (g - 1) % num_gpus )) ${cmd} "${dir}/log/extract.${g}.log" \
      python "${here}/extract_embedding.py" \
        --use-gpu=yes --min-chunk-size=${min_chunk_size} --chunk-size=${chunk_size} \
        --feature-rspecifier="`echo ${feature_rspecifier} | sed s/JOB/${g}/g`" \
        --vector-wspecifier="| copy-vector ark:- ark,scp:${dir}/xvector.${g}.ark,${dir}/xvector.${g}.scp" \
        --model-dir="${model_dir}" || exit 1 &
  done
  wait
fi

if [ ${stage} -le 1 ]; then
  echo "${0}: combining xvectors across jobs"
  for j in $(seq ${nj}); do cat ${dir}/xvector.${j}.scp; done > ${dir}/xvector.scp || exit 1;
fi

if [ ${stage} -le 2 ]; then
  # Average the utterance-level xvectors to get speaker-level xvectors.
  echo "${0}: computing mean of xvectors for each speaker"
  run.pl ${dir}/log/speaker_mean.log \
    ivector-mean ark:${data}/spk2utt scp:${dir}/xvector.scp \
    ark,scp:${dir}/spk_xvector.ark,${dir}/spk_xvector.scp ark,t:${dir}/num_utts.ark || exit 1;
fi
