"""Shadow of kernel/go_model.py."""
from igcn_b200.go_net import Gene_ontology_network  # noqa: F401
