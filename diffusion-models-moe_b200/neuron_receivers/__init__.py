"""Drop-in `neuron_receivers` package for the hot-path receivers of
ruchikachavhan/diffusion-models-moe, computing on libmoe_b200.so (B200 / sm_100a).
Same class names and call surface as the reference's neuron_receivers/__init__.py:1-19
(plus the SURVEY 8f row-1 receivers built on the same kernels: SparsityMeasure, GetExperts, AddExperts, Wanda)."""
from neuron_receivers.base_receiver import BaseNeuronReceiver
from neuron_receivers.frequency_measure import FrequencyMeasure
from neuron_receivers.moefy import MOEFy
from neuron_receivers.predictivity import NeuronPredictivity
from neuron_receivers.remove_skilled_experts import RemoveExperts
from neuron_receivers.remove_skilled_neurons import RemoveNeurons
from neuron_receivers.expert_activation import ExpertPredictivity
from neuron_receivers.remove_wanda_neurons_fast import WandaRemoveNeuronsFast
from neuron_receivers.multi_concept_remover import MultiConceptRemoverWanda
from neuron_receivers.sparsity_measure import SparsityMeasure
from neuron_receivers.get_experts import GetExperts
from neuron_receivers.add_skilled_experts import AddExperts
from neuron_receivers.wanda_receiver import Wanda

__all__ = ["BaseNeuronReceiver", "FrequencyMeasure", "MOEFy", "NeuronPredictivity", "RemoveExperts",
           "RemoveNeurons", "ExpertPredictivity", "WandaRemoveNeuronsFast", "MultiConceptRemoverWanda", "SparsityMeasure",
           "GetExperts", "AddExperts", "Wanda"]
