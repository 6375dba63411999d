"""Model zoo for the BASELINE.json configs (Llama-2, GPT-3, Mixtral-MoE, ResNet-50, MNIST MLP)."""
from .llama import LlamaConfig, LlamaDecoderLayer, LlamaForCausalLM, LlamaModel, LlamaPretrainingCriterion, llama2_13b, llama_tiny  # noqa: F401
