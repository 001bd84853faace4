"""Host-side mirror of the reference's Python interface for the post-backbone hot path.

Module paths, class / function names, argument order, defaults and return types follow the
reference (file:line cited in each docstring) so its train.py / eval.py call sites keep working when
`rrnet_b200.host` is put in front of the reference on sys.path:

    reference module                         mirror
    models/rrnet.py                          rrnet_b200.host.models.rrnet            (RRNet)
    detectors/fasterrcnn_detector.py         rrnet_b200.host.detectors.fasterrcnn_detector
    operators/rrnet_operator.py              rrnet_b200.host.operators.rrnet_operator (RRNetOperator)
    modules/loss/{focalloss,functional}.py   rrnet_b200.host.modules.loss.*          (FocalLossHM, focal_loss_for_hm)
    datasets/transforms/{functional,transforms}.py  rrnet_b200.host.datasets.transforms.*   (to_heatmap, ToHeatmap)
    ext/nms/nms_wrapper.py, ext/nms/nms/*    rrnet_b200.host.ext.nms.*               (nms, soft_nms, gpu_nms, cpu_nms, ...)

Everything computes through librrnet_b200.so (rrnet_b200.ops); there is no CPU fallback.  The
stage-1 convolution heads and the backbone are not part of this path: RRNet takes them as modules.
`rrnet_b200.host.sharding` holds the image-sharded multi-GPU eval driver (NCCL all-gather of detections).
"""
