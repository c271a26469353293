"""Util functions for Uniform Quantization — the graph-editing surface of the reference
(/root/reference/learners/uniform_quantization/utils.py:31-306).

The reference splices TF quantization sub-graphs with tf.contrib.graph_editor; here the same
search / insert calls MARK the ops, and engine.Executor lowers the marks to the fused CUDA kernels
(pf_uq_weight_minmax/scales/quant, pf_bn_apply(+range), pf_uq_act_quant)."""


def prefix_filter(prefix):
    """filter out the variable_scope"""
    ind = prefix.index('/')
    return prefix[ind + 1:]


class UniformQuantization:
    # pylint: disable=too-many-instance-attributes
    """ Class of uniform quantization """

    def __init__(self, sess, bucket_size=0, use_buckets=False, bucket_type='split'):
        self.sess = sess                      # the graph being edited (stands in for sess.graph)
        self.use_buckets = use_buckets
        self.bucket_size = bucket_size
        self.bucket_type = bucket_type
        self.matmul_ops = []
        self.activation_ops = []
        self.quantized_matmul_ops = []
        self.quantized_activation_ops = []
        self.bucket_storage = 0  # bits
        self.weight_bits = []
        self.activation_bits = []
        self.__safe_check()
        self.support_act_types = ['Relu', 'Relu6', 'Crelu', 'Elu', 'Selu', 'Softplus',
                                  'Softsign', 'Sigmoid', 'Tanh']
        self.support_mul_types = ['Conv2D', 'MatMul', 'DepthwiseConv2dNative']

    def insert_quant_op_for_activations(self, act_bit_dict):
        """act_bit_dict: (key: act_op_name, value: act_bits).  Only Relu / Relu6 have kernels."""
        for op in self.activation_ops:
            if op.type in ('Relu', 'Relu6'):
                self.quantized_activation_ops.append(op)
                self.activation_bits.append(int(act_bit_dict[op.name]))
            elif op.type in self.support_act_types:
                raise NotImplementedError("The activation_fn needs to include %s manually" % op.type)
            else:
                raise ValueError("Unknown activation mode, you may add it manually here")

    def insert_quant_op_for_weights(self, w_bit_dict):
        """w_bit_dict: (key: matmul_op_name, value: quant_bits)"""
        from ... import ops as _ops
        for op in self.matmul_ops:
            if op.type not in self.support_mul_types:
                raise NotImplementedError("Unrecognied Mul op, try to add it into matmul_typs for quantization")
            self.quantized_matmul_ops.append(op)
            self.weight_bits.append(int(w_bit_dict[op.name]))
            if self.use_buckets:
                ncols, _ = _ops.uq_bucket_layout(op.vars['kernel'].shape, True, self.bucket_type, self.bucket_size)
                self.__updt_bucket_storage(ncols)

    def search_matmul_op(self, quantize_all_layers):
        """ search matmul or Conv2D operations in graph for quantization"""
        is_student_fn = lambda x: 'distilled' not in x.name
        for op in self.sess.get_operations():
            if op.type in self.support_mul_types and is_student_fn(op):
                self.matmul_ops.append(op)
        if not quantize_all_layers:
            self.matmul_ops = self.matmul_ops[1:-1]  # remain full precision for first and last layer
        return self.matmul_ops

    def search_activation_op(self):
        """ search activation operation in graph for quantization """
        is_student_fn = lambda x: 'distilled' not in x.name
        for op in self.sess.get_operations():
            if op.type in self.support_act_types and is_student_fn(op):
                self.activation_ops.append(op)
        return self.activation_ops

    def weight_quant_spec(self):
        if not self.quantized_matmul_ops:
            return None
        return dict(kind='uniform', ops=self.quantized_matmul_ops, bits=self.weight_bits,
                    use_buckets=self.use_buckets, bucket_type=self.bucket_type, bucket_size=self.bucket_size)

    def act_quant_spec(self):
        if not self.quantized_activation_ops:
            return None
        return dict(ops=self.quantized_activation_ops, bits=self.activation_bits)

    def __safe_check(self):
        if self.bucket_size < 0:
            raise ValueError("Bucket size must be a postive integer")
        if self.bucket_type != 'split' and self.bucket_type != 'channel':
            raise ValueError("Unrecognized bucket type, must be 'weight' or 'channel'.")

    def __updt_bucket_storage(self, bucket_num):
        self.bucket_storage += bucket_num * 32 * 2  # both alpha and beta, so *2
