from .meta_arch.build import META_ARCH_REGISTRY, build_model
