"""TaskFunc hooks used inside the training step (task/sseg/func.py:134-253).  Validation-time
``metrics`` / ``visualize`` are out of scope for this round."""


def task_func():
    return SemanticSegmentationFunc


class SemanticSegmentationFunc:
    METRIC_STR = 'metric'

    def __init__(self, args):
        self.args = args

    def sslcct_ad_in_channels(self):
        return 2048

    def sslcct_ad_out_channels(self):
        return self.args.num_classes

    def sslcct_ad_upsample_scale(self):
        return 8

    def sslgct_fd_in_channels(self):
        return self.args.num_classes + 3

    def ssladv_fcd_in_channels(self):
        return self.args.num_classes
