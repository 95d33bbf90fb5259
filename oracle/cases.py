"""TEST INFRASTRUCTURE ONLY -- the seeded small cases shared by golden generation and the tests."""
from centermask2_b200.config import get_cfg, lite_overrides

# name -> (cfg overrides, [(h, w), ...] image sizes, weight seed, image seed, candidates/level target)
CASES = {
    "v19_two_images": (["MODEL.VOVNET.CONV_BODY", "V-19-eSE"], [(96, 128), (80, 120)], 11, 21, 300),
    "v39_one_image": ([], [(128, 160)], 12, 22, 400),
    "v19_lite": (lite_overrides(), [(64, 96), (64, 96), (60, 90)], 13, 23, 200),
    "v19_empty": (["MODEL.VOVNET.CONV_BODY", "V-19-eSE"], [(64, 64)], 14, 24, 0),
    "v19_crowded": (["MODEL.VOVNET.CONV_BODY", "V-19-eSE", "MODEL.FCOS.POST_NMS_TOPK_TEST", 100],
                    [(128, 128)], 15, 25, 3000),
    "v99_one_image": (["MODEL.VOVNET.CONV_BODY", "V-99-eSE"], [(64, 96)], 16, 26, 300),
    # depthwise body (vovnet.py:30-38, :110-130): dw 3x3 + pw 1x1 units, 1x1 reduction in stages 3-5, stem 64/64/64
    "v19_slim_dw": (["MODEL.VOVNET.CONV_BODY", "V-19-slim-dw-eSE"], [(96, 128), (72, 100)], 17, 27, 300),
    # keypoint branch (center_heads.py:358-383,520-553; keypoint_head.py:95-222): narrow tower to keep the case small
    "v19_keypoints": (["MODEL.VOVNET.CONV_BODY", "V-19-eSE", "MODEL.KEYPOINT_ON", True,
                       "MODEL.ROI_KEYPOINT_HEAD.IN_FEATURES", ["p3", "p4", "p5"],
                       "MODEL.ROI_KEYPOINT_HEAD.CONV_DIMS", (128, 128, 64), "MODEL.FCOS.POST_NMS_TOPK_TEST", 20],
                      [(96, 128), (80, 120)], 18, 28, 300),
    # ---- variants of SURVEY 8f-4, pinned against the unmodified reference (lean goldens: raw + post only)
    # FPN Eq. 1 level rule (pooler.py:121-152): a 512x640 image so that boxes reach the canonical sizes of P4 / P5
    "v19_area": (["MODEL.VOVNET.CONV_BODY", "V-19-eSE", "MODEL.ROI_MASK_HEAD.ASSIGN_CRITERION", "area",
                  "MODEL.FCOS.POST_NMS_TOPK_TEST", 20], [(512, 640)], 31, 41, 300, {"lean": True}),
    # centerness folded into the score BEFORE the 0.05 threshold (fcos_outputs.py:412-413)
    "v19_ctr_thresh": (["MODEL.VOVNET.CONV_BODY", "V-19-eSE", "MODEL.FCOS.THRESH_WITH_CTR", True], [(96, 128), (80, 120)], 32, 42, 400,
                       {"lean": True}),
    # LastLevelP6 (fpn.py:38-53; vovnet.py:543-544): one extra level, four FCOS levels
    "v19_p6": (["MODEL.VOVNET.CONV_BODY", "V-19-eSE", "MODEL.FCOS.TOP_LEVELS", 1, "MODEL.FCOS.IN_FEATURES", ["p3", "p4", "p5", "p6"],
                "MODEL.FCOS.FPN_STRIDES", [8, 16, 32, 64]], [(96, 128)], 33, 43, 300, {"lean": True}),
    # the fork's tensor-in / tuple-out meta-architecture (modified_class.py:27-40), run through the reference's own class
    "v19_tensor_in": (["MODEL.VOVNET.CONV_BODY", "V-19-eSE"], [(96, 128)], 34, 44, 300, {"lean": True, "tensor_in": True}),
}


def case_flags(name):
    c = CASES[name]
    return c[5] if len(c) > 5 else {}


def case_cfg(name):
    return get_cfg("centermask_V_39_eSE_FPN.yaml", CASES[name][0])
