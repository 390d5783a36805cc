"""``PseudoGenerator`` (reference: ubteacher/modeling/pseudo_generator.py:7-131) on device-resident
detections: a second NMS criterion over the teacher's dense outputs and the threshold-scatter that turns
detections into pseudo-label sets, with no host synchronisation."""
from .. import ops
from .fcos.fcos_outputs import BoxSet, FCOSOutputs


class PseudoGenerator:
    def __init__(self, cfg):
        self.fcos_output = FCOSOutputs(cfg)     # stays in train mode like the reference's private copy (A.3 #7)

    def nms_from_dense(self, raw_output, nms_method, scales=None):
        assert nms_method in ["cls", "ctr", "cls_n_ctr", "cls_n_loc"]
        scales = raw_output["scales"] if scales is None else scales
        return self.fcos_output.predict_proposals(raw_output, scales, nms_method)

    def process_pseudo_label(self, proposals_rpn_unsup_k, cur_threshold, proposal_type, psedo_label_method=""):
        if psedo_label_method == "thresholding":
            out = self.threshold_bbox(proposals_rpn_unsup_k, thres=cur_threshold, proposal_type=proposal_type)
        elif psedo_label_method == "thresholding_cls_ctr":
            out = self.threshold_cls_ctr_bbox(proposals_rpn_unsup_k, thres=cur_threshold)
        else:
            raise ValueError("Unkown pseudo label boxes methods")
        return out, out.counts.float().mean()

    @staticmethod
    def _boxset(o):
        return BoxSet(o["pred_boxes"], o["pred_classes"], o["count"], o["reg_pred_std"], o["scores"],
                      {"centerness": o["centerness"], "cls_confid": o["cls_confid"]})

    def threshold_bbox(self, dets, thres=0.7, proposal_type="roih"):
        if proposal_type != "roih":
            raise NotImplementedError("proposal_type 'rpn' is the R-CNN trainer's path")
        return self._boxset(ops.threshold_scatter(dets, 0, float(thres)))

    def threshold_cls_ctr_bbox(self, dets, thres=(0.5, 0.5)):
        return self._boxset(ops.threshold_scatter(dets, 1, float(thres[0]), float(thres[1])))
