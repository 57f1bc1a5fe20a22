"""Drop-in replacements for the reference's `model/` package (model.model, model.video_transformer, model.loss,
model.metric): same class names, constructor arguments, forward signatures and state_dict keys."""
