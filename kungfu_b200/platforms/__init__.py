"""Platform adapters (reference: srcs/go/platforms/modelarts)."""
