pub mod pca;
