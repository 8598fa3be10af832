"""Drop-in `neuron_receivers` package for the hot-path receivers of
ruchikachavhan/diffusion-models-moe, computing on libmoe_b200.so (B200 / sm_100a).
Same class names and call surface as the reference's neuron_receivers/__init__.py:1-19
(the receivers outside the MoEfied-FFN hot path are not part of this package)."""
from neuron_receivers.base_receiver import BaseNeuronReceiver
from neuron_receivers.frequency_measure import FrequencyMeasure
from neuron_receivers.moefy import MOEFy
from neuron_receivers.predictivity import NeuronPredictivity
from neuron_receivers.remove_skilled_experts import RemoveExperts
from neuron_receivers.remove_skilled_neurons import RemoveNeurons
from neuron_receivers.expert_activation import ExpertPredictivity
from neuron_receivers.remove_wanda_neurons_fast import WandaRemoveNeuronsFast
from neuron_receivers.multi_concept_remover import MultiConceptRemoverWanda

__all__ = ["BaseNeuronReceiver", "FrequencyMeasure", "MOEFy", "NeuronPredictivity", "RemoveExperts",
           "RemoveNeurons", "ExpertPredictivity", "WandaRemoveNeuronsFast", "MultiConceptRemoverWanda"]
