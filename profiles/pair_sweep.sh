for shape in "0 3 1024 8 8 256 0 256" "0 3 1024 8 8 256 256 128" "0 3 1024 8 8 128 0 256" "2 4 1024 8 8 128 0 128" "0 3 1024 4 4 256 0 256" "0 3 1024 8 8 128 0 128" "1 3 1024 16 16 128 0 128" "0 1 1024 8 8 256 256 128"; do
  for pm in 0 1; do echo -n "pair=$pm "; DBS_CONV_PAIR=$pm timeout 90 python profiles/conv_micro.py $shape 2>&1 | tail -1; done
done
