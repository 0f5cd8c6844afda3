"""Drop-in module path of the reference (`from functions.functions_RESNET50_Truncate_Gram_Attention import ...`,
train_best_RESNET50_Truncate_gram_attention.py:12-18, test_RESNET50_Truncate_gram_attention.py:11-19). The functions
live in heuristique_style_transfer_code_b200/functions.py."""
from heuristique_style_transfer_code_b200.functions import (  # noqa: F401
    create_onpick_function,
    denormalize,
    evaluate_model,
    evaluate_model_test,
    load_hyperparameters,
    load_model,
    load_model_weights,
    perform_tsne,
    plot_tsne_interactive,
    run_camera,
    save_model_weights,
    set_parameter_requires_grad,
    style_transfer,
    train_model,
)
