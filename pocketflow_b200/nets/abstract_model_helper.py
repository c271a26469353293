"""Abstract class for model helpers — the plugin surface of the reference, kept source-compatible
(/root/reference/nets/abstract_model_helper.py:22-149).  Where the reference passes TF tensors,
`pocketflow_b200.graph.Tensor` nodes are passed; where it passes `sess`, the learner's executor."""
from abc import ABC
from abc import abstractmethod


class AbstractModelHelper(ABC):
    """A model helper defines: 1. data input pipelines, 2. the network's forward pass at training
    and evaluation, 3. the loss function (and extra evaluation metrics)."""

    def __init__(self, data_format, forward_w_labels=False):
        self.data_format = data_format
        self.forward_w_labels = forward_w_labels

    @abstractmethod
    def build_dataset_train(self, enbl_trn_val_split):
        """Returns an iterator (or a pair with enbl_trn_val_split) over training mini-batches."""

    @abstractmethod
    def build_dataset_eval(self):
        """Returns an iterator over evaluation mini-batches."""

    @abstractmethod
    def forward_train(self, inputs, labels=None):
        """Forward computation at training: inputs -> outputs (graph nodes)."""

    @abstractmethod
    def forward_eval(self, inputs):
        """Forward computation at evaluation."""

    @abstractmethod
    def calc_loss(self, labels, outputs, trainable_vars):
        """Returns (loss, metrics dict)."""

    @abstractmethod
    def setup_lrn_rate(self, global_step):
        """Returns (lrn_rate, nb_iters)."""

    def warm_start(self, sess):
        """Initialize the model for warm-start."""

    def dump_n_eval(self, outputs, action):
        """Dump the model's outputs to files and evaluate ('init' | 'dump' | 'eval')."""

    @property
    @abstractmethod
    def model_name(self):
        """Model's name."""

    @property
    @abstractmethod
    def dataset_name(self):
        """Dataset's name."""
