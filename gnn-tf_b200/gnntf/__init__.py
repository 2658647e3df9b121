"""gnntf — drop-in surface of MKLab-ITI/gnn-tf's sparse adjacency propagation path, running on
hand-written sm_100a kernels (``libgnntf_b200.so``) behind a C-ABI.  See DESIGN.md.

Exports mirror ``gnntf/__init__.py:1-2`` of the reference for the path in scope."""
from .graph_manipulation import adj2graph, create_nx_graph, csr2adj, edges2adj, graph2adj, graph2indices
from .measures import acc, set_seed
from .nn import (Activation, Concatenate, Dense, Dropout, Layer, Layered, Predictor, Trainable, VariableGenerator,
                 WrappedVariable)
from .gnn import (APPNP, GCN, GCNII, GCNIILayer, GCNIISpectralPreservingLayer, GCNLayer, GCNSpectralPreservingLayer, GNN,
                  NGCF, NGCFLayer, NodeClassification, PPRIteration, Structural)
from .ops import (appnp_propagate, appnp_propagate_host, appnp_propagate_host_batched, appnp_step, bias_act_dropout, node_cross_entropy,
                  sparse_dense_matmul)
from .sparse import NormalizedAdjacency, SparseAdjacency

__version__ = "0.1.0"
