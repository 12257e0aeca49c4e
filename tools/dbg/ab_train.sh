for i in 1 2; do for v in old b200; do IIC_LIB=$PWD/ai-interior-image-classifier_b200/_lib/libiic_$v.so python tools/train_ab.py 2>/dev/null | tail -1; done; done
